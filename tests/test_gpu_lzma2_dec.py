"""GPU parity tests of the LZMA2 (method 21) decoder through the C ABI: bit-exact against the plain-C oracle and the
original data on the committed golden streams (reference regression archive, reference encoders, liblzma), on streams
made here by the reference's encoders (oracle/_ref) with many independent blocks, and on corrupted input."""
import hashlib
import json
import lzma
import os
import random

import pytest

import helpers

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
IDX = json.load(open(os.path.join(GOLDEN, "lzma2.json")))


@pytest.mark.parametrize("name", sorted(IDX))
def test_golden_streams(codec, name):
    meta = IDX[name]
    comp = open(os.path.join(GOLDEN, name), "rb").read()
    out = codec.lzma2_decompress(comp, meta["dict_prop"])
    assert len(out) == meta["size"] and hashlib.sha256(out).hexdigest() == meta["sha256"]


def test_liblzma_streams(codec, pkg):
    for seed, n, preset, lc, lp, pb in [(1, 50_000, 0, 3, 0, 2), (2, 200_000, 4, 0, 0, 0), (3, 90_000, 9, 4, 0, 4), (4, 3_000_000, 1, 1, 2, 3), (5, 70_000, 6, 0, 4, 2)]:
        data = pkg.corpus.g2(n, seed=seed).tobytes()
        comp = lzma.compress(data, format=lzma.FORMAT_RAW, filters=[dict(id=lzma.FILTER_LZMA2, preset=preset, dict_size=1 << 16, lc=lc, lp=lp, pb=pb)])
        assert codec.lzma2_decompress(comp, 8) == data


@pytest.mark.skipif(not helpers.ref_lzma_available(), reason="oracle/_ref/libref_lzma.so not built")
def test_reference_encoder_streams(codec, pkg):
    ins = helpers.sample_inputs(pkg, big=True)
    for name, data in ins.items():
        for mk in ("fl2", "blocks"):
            if mk == "fl2":
                prop, comp = helpers.ref_fl2_compress(data, 3, threads=2)
            else:       # stock encoder, independent 1 MiB blocks (Lzma2Enc.c:241-330): every block starts with a dictionary reset
                prop, comp = helpers.ref_lzma2_compress(data, 1, dict_size=1 << 18, block_size=1 << 20, threads=4)
            size, nblk, used = codec.lzma2_stream_info(comp)
            assert size == len(data) and comp[used - 1] == 0
            if mk == "blocks" and len(data) > (2 << 20):
                assert nblk >= len(data) >> 20
            out = codec.lzma2_decompress(comp, prop)
            assert out == data, (name, mk)
            assert helpers.oracle_lzma2_decompress(comp, len(data), prop)[0] == out


def test_errors_and_corruption(codec, pkg):
    comp = open(os.path.join(GOLDEN, "lzma2_fl2_g2_100k.bin"), "rb").read()
    meta = IDX["lzma2_fl2_g2_100k.bin"]
    with pytest.raises(pkg.B200zError) as e:
        codec.lzma2_decompress(comp[:1000], meta["dict_prop"], max_size=meta["size"])
    assert e.value.code == -5
    with pytest.raises(pkg.B200zError) as e:
        codec.lzma2_decompress(comp, meta["dict_prop"], max_size=50_000)
    assert e.value.code == -4
    with pytest.raises(pkg.B200zError) as e:
        codec.lzma2_decompress(b"\x80" + comp[1:], meta["dict_prop"], max_size=meta["size"])       # no dictionary reset
    assert e.value.code == -5
    with pytest.raises(pkg.B200zError):
        codec.lzma2_decompress(comp, 41, max_size=meta["size"])
    # dictionary smaller than the distances used -> data error, as in the reference (LzmaDec.c checkDicSize)
    rng = random.Random(5)
    for _ in range(40):
        bad = bytearray(comp); k = rng.randrange(6, len(comp) - 8); bad[k] ^= 1 << rng.randrange(8)
        try:
            want = helpers.oracle_lzma2_decompress(bytes(bad), meta["size"], meta["dict_prop"])[0]
        except ValueError:
            want = None
        try:
            got = codec.lzma2_decompress(bytes(bad), meta["dict_prop"], max_size=meta["size"])
        except pkg.B200zError as ex:
            assert ex.code in (-5, -4)        # -4: a corrupted chunk header claims more output than the caller allows
            got = None
        assert got == want
