"""Pins oracle/lzma2_dec_oracle.c (CPU): golden LZMA2 streams from the reference's own regression archive, from the
reference's two encoders and from liblzma; plus agreement with the reference decoder when oracle/_ref is present."""
import hashlib
import json
import lzma
import os

import pytest

import helpers as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
IDX = json.load(open(os.path.join(GOLD, "lzma2.json")))


@pytest.mark.parametrize("name", sorted(IDX))
def test_golden_streams(name):
    meta = IDX[name]
    comp = open(os.path.join(GOLD, name), "rb").read()
    out, used = H.oracle_lzma2_decompress(comp, meta["size"], meta["dict_prop"])
    assert len(out) == meta["size"] and hashlib.sha256(out).hexdigest() == meta["sha256"]
    assert comp[used - 1] == 0 and used <= len(comp)          # stops on the end marker (FL2 appends a hash after it)
    if H.ref_lzma_available():
        r, rused = H.ref_lzma2_decompress(comp, meta["size"], meta["dict_prop"])
        assert r == out and rused == used


def test_liblzma_streams(pkg):
    for seed, n, preset, lc, lp, pb in [(1, 50_000, 0, 3, 0, 2), (2, 200_000, 4, 0, 0, 0), (3, 90_000, 9, 4, 0, 4), (4, 3_000_000, 1, 1, 2, 3)]:
        data = pkg.corpus.g2(n, seed=seed).tobytes()
        comp = lzma.compress(data, format=lzma.FORMAT_RAW, filters=[dict(id=lzma.FILTER_LZMA2, preset=preset, dict_size=1 << 16, lc=lc, lp=lp, pb=pb)])
        out, used = H.oracle_lzma2_decompress(comp, n, 8)
        assert out == data and used == len(comp)


def test_errors():
    comp = open(os.path.join(GOLD, "lzma2_fl2_g2_100k.bin"), "rb").read()
    with pytest.raises(ValueError):
        H.oracle_lzma2_decompress(comp[:1000], 100_000, 10)              # truncated
    with pytest.raises(ValueError):
        H.oracle_lzma2_decompress(comp, 50_000, 10)                      # destination too small
    with pytest.raises(ValueError):
        H.oracle_lzma2_decompress(b"\x80" + comp[1:], 100_000, 10)       # first chunk without dictionary reset
    bad = bytearray(comp); bad[3000] ^= 0x55
    try:
        out, _ = H.oracle_lzma2_decompress(bytes(bad), 100_000, 10)      # corruption: an error or different bytes, never a crash
        assert hashlib.sha256(out).hexdigest() != IDX["lzma2_fl2_g2_100k.bin"]["sha256"]
    except ValueError:
        pass
