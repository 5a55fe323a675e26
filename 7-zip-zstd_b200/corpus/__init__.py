"""Seeded synthetic corpora of the BASELINE.json configs (C generator: corpus/g2gen.c)."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
G2_SEED = 20260922


def _load():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libb200z_corpus.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing -- run 7-zip-zstd_b200/build.sh")
        _lib = ctypes.CDLL(path)
        _lib.b200z_corpus_g2.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int]
        _lib.b200z_corpus_class.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64]
    return _lib


def g2_into(ptr, nbytes, seed=G2_SEED, offset=0, threads=None):
    """Fill [ptr, ptr+nbytes) with G2 text (SURVEY.md 8(d) cfg2 shape); offset must be a multiple of 1 MiB."""
    threads = threads or min(64, os.cpu_count() or 1)
    rc = _load().b200z_corpus_g2(seed, offset, ptr, nbytes, threads)
    if rc:
        raise ValueError("offset must be a multiple of 1 MiB")


def g2(nbytes, seed=G2_SEED, offset=0, threads=None) -> np.ndarray:
    buf = np.empty(nbytes, dtype=np.uint8)
    if nbytes:
        g2_into(buf.ctypes.data, nbytes, seed, offset, threads)
    return buf


def entropy_class(cls, nbytes, seed=5) -> np.ndarray:
    """cfg5 file classes: 0 text, 1 8-bit noise, 2 16-symbol skewed, 3 tiled 4 KiB with 1 % noise."""
    buf = np.empty(nbytes, dtype=np.uint8)
    if nbytes:
        _load().b200z_corpus_class(seed, cls, buf.ctypes.data, nbytes)
    return buf


def inject_far_copies(buf, every=64 << 20, span=(1 << 20, 4 << 20), back=128 << 20, mutate=0.001, seed=3) -> int:
    """cfg3: long-range redundancy written into `buf` in place -- at every multiple of `every`, a span of
    span[0]..span[1] bytes copied from a uniformly random earlier position (at most `back` bytes back), `mutate` of its bytes
    changed.  Returns the bytes planted."""
    rng = np.random.default_rng(seed)
    n = buf.size
    planted = 0
    for at in range(every, n, every):
        ln = min(int(rng.integers(span[0], span[1])), n - at)
        lo = 0 if back is None else max(0, at - back)
        if at - ln <= lo:
            continue
        srcp = int(rng.integers(lo, at - ln))
        seg = buf[srcp:srcp + ln].copy()
        m = rng.random(ln) < mutate
        seg[m] = rng.integers(0, 256, int(m.sum()), dtype=np.uint8)
        buf[at:at + ln] = seg
        planted += ln
    return planted


def g3(nbytes, seed=3) -> np.ndarray:
    """cfg3's input: G2 text + inject_far_copies with the recipe's constants"""
    buf = g2(nbytes)
    inject_far_copies(buf, seed=seed)
    return buf
