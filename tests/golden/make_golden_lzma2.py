"""Regenerates the LZMA2 (method 21) fixtures of tests/golden/ from the reference (build container only).

test.txt.lzma2   the packed stream of the reference's own tests/regr-arc/test.txt.7z (bytes 32..32+NextHeaderOffset;
                 the archive's folder record names coder 21 with property byte 0x10), whose plaintext SHA-256 the
                 reference's regression suite pins (tests/regression.test).
lzma2_*.bin      outputs of the reference's encoders compiled into oracle/_ref/libref_lzma.so (Fast-LZMA2: FL2_compressMt,
                 stock: Lzma2Enc_Encode2) and of liblzma (python `lzma`, an independent encoder) on seeded inputs."""
import hashlib
import json
import lzma
import os
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(HERE))
import __graft_entry__ as ge  # noqa: E402
import helpers  # noqa: E402

pkg = ge.load_package()
g2, cls = pkg.corpus.g2, pkg.corpus.entropy_class
arc = open("/root/reference/tests/regr-arc/test.txt.7z", "rb").read()
nho, = struct.unpack("<Q", arc[12:20])
assert arc[32 + nho:].find(bytes([0x21, 0x21, 0x01, 0x10])) > 0        # coder id 21, 1 property byte, 0x10
open(os.path.join(HERE, "test.txt.lzma2"), "wb").write(arc[32:32 + nho])

idx = {"test.txt.lzma2": {"size": 1000000, "dict_prop": 0x10, "by": "tests/regr-arc/test.txt.7z",
                          "sha256": hashlib.sha256(b"TEST\n" + b" " * 999990 + b"\nEND.").hexdigest()}}


def py_raw(data, **f):
    return lzma.compress(data, format=lzma.FORMAT_RAW, filters=[dict(id=lzma.FILTER_LZMA2, **f)])


cases = {
    "fl2_g2_100k": (g2(100_000).tobytes(), lambda d: helpers.ref_fl2_compress(d, 5), "FL2 L5"),
    "fl2_zeros": (bytes(500_000), lambda d: helpers.ref_fl2_compress(d, 3), "FL2 L3"),
    "fl2_mt_g2_600k": (g2(600_000, seed=4).tobytes()[::3] * 1, lambda d: helpers.ref_fl2_compress(d, 1, threads=3), "FL2 L1 3 threads (slices reset state)"),
    "lzma2_tile_blocks": (cls(3, 900_000).tobytes(), lambda d: helpers.ref_lzma2_compress(d, 1, dict_size=1 << 16, lc=2, lp=1, pb=1, block_size=1 << 18, threads=2),
                          "Lzma2Enc L1 dict 64K lc2 lp1 pb1, 256 KiB independent blocks"),
    "lzma2_noise_70k": (cls(1, 70_000).tobytes(), lambda d: helpers.ref_lzma2_compress(d, 5), "Lzma2Enc L5: uncompressed chunks"),
    "lzma2_lc0_lp4": (g2(60_000, seed=2).tobytes(), lambda d: helpers.ref_lzma2_compress(d, 9, lc=0, lp=4, pb=4), "Lzma2Enc L9 lc0 lp4 pb4"),
    "py_g2_128k1": (g2(131073).tobytes(), lambda d: (18, py_raw(d, preset=6, dict_size=1 << 20)), "liblzma preset 6"),
    "py_empty": (b"", lambda d: (18, py_raw(d, preset=1, dict_size=1 << 20)), "liblzma, empty input"),
}
for name, (data, fn, by) in cases.items():
    prop, comp = fn(data)
    f = f"lzma2_{name}.bin"
    open(os.path.join(HERE, f), "wb").write(comp)
    ref_out, _ = helpers.ref_lzma2_decompress(comp, len(data), prop)
    assert ref_out == data
    idx[f] = {"size": len(data), "dict_prop": prop, "by": by, "sha256": hashlib.sha256(data).hexdigest()}
json.dump(idx, open(os.path.join(HERE, "lzma2.json"), "w"), indent=1, sort_keys=True)
print({k: os.path.getsize(os.path.join(HERE, k)) for k in idx})
