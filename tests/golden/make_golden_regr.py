"""Extracts the packed streams of the reference's own regression archives (tests/regr-arc/*.7z, the fixtures behind
tests/regression.test:31-89,177-213) for the two methods on the hot path.  The archives' headers (plain for test.txt*.7z,
LZMA-encoded for test-sol*.7z) give: one folder, one coder, pack stream at offset 32; sizes below are the NextHeaderOffset /
PackPos fields read from those headers.  Expected payloads are the ones regression.test builds (:181-182)."""
import hashlib
import json
import os
import struct

HERE = os.path.dirname(os.path.abspath(__file__))
R = "/root/reference/tests/regr-arc/"
txt = b"TEST\n" + b" " * 999990 + b"\nEND."
sol = txt + b"1234\n" + b" " * 99990 + b"\n5678"
# archive -> (output name, packed bytes, payload, method, LZMA2 dictionary property)
cases = {
    "test.txt.zstd.7z": ("regr_test.txt.l17.zst", None, txt, "zstd", None),            # coder 04F71101, props 01 05 11 00 00 (level 17)
    "test-sol.zstd.7z": ("regr_test-sol.l17.zst", 89, sol, "zstd", None),              # solid folder of both files
    "test-sol.zstd.max.7z": ("regr_test-sol.max.zst", 89, sol, "zstd", None),          # ZSTD:max
    "test-sol.7z": ("regr_test-sol.lzma2", 261, sol, "lzma2", 17),                     # LZMA2:1536k
}
idx = {}
for arc, (out, n, payload, method, prop) in cases.items():
    b = open(R + arc, "rb").read()
    if n is None:
        n, = struct.unpack("<Q", b[12:20])                                              # plain header: pack stream ends where the header starts
    open(os.path.join(HERE, out), "wb").write(b[32:32 + n])
    idx[out] = {"from": "tests/regr-arc/" + arc, "method": method, "size": len(payload), "sha256": hashlib.sha256(payload).hexdigest()}
    if prop is not None:
        idx[out]["dict_prop"] = prop
json.dump(idx, open(os.path.join(HERE, "regr.json"), "w"), indent=1, sort_keys=True)
print({k: os.path.getsize(os.path.join(HERE, k)) for k in idx})
