"""Regenerates tests/golden/ from the reference (run in the build container, where /root/reference
and oracle/_ref exist).  The .zstd golden vector is the reference's own tests/regr-arc/test.txt.zstd;
frames_*.zst are outputs of the reference encoder (oracle/_ref/libref_zstd.so) on seeded inputs."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(HERE))
import __graft_entry__ as ge  # noqa: E402
import helpers  # noqa: E402

pkg = ge.load_package()
shutil.copy("/root/reference/tests/regr-arc/test.txt.zstd", os.path.join(HERE, "test.txt.zstd"))
open(os.path.join(HERE, "test.txt.sha256"), "w").write(
    hashlib.sha256(b"TEST\n" + b" " * 999990 + b"\nEND.").hexdigest() + "\n")
idx = {}
cases = {
    "g2_l3": (pkg.corpus.g2(300_000).tobytes(), dict(level=3)),
    "g2_l19_ck": (pkg.corpus.g2(200_000, seed=9).tobytes(), dict(level=19, checksum=1)),
    "g2_fast5": (pkg.corpus.g2(150_000, seed=3).tobytes(), dict(level=-5)),
    "tile_l9": (pkg.corpus.entropy_class(3, 200_000).tobytes(), dict(level=9)),
    "skew_l1": (pkg.corpus.entropy_class(2, 120_000).tobytes(), dict(level=1, checksum=1)),
    "mixed_l6": (pkg.corpus.g2(70_000).tobytes() + bytes(50_000) + pkg.corpus.entropy_class(1, 30_000).tobytes(), dict(level=6)),
}
for name, (data, kw) in cases.items():
    comp = helpers.ref_compress(data, **kw)
    fn = f"frames_{name}.zst"
    open(os.path.join(HERE, fn), "wb").write(comp)
    idx[fn] = {"size": len(data), "sha256": hashlib.sha256(data).hexdigest(), "params": kw}
json.dump(idx, open(os.path.join(HERE, "frames.json"), "w"), indent=1, sort_keys=True)
print({k: os.path.getsize(os.path.join(HERE, k)) for k in idx})
