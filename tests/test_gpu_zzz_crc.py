"""GPU: CRC32 / CRC-64 of host buffers through the C ABI (csrc/b2z_crc.cu) against zlib and the oracle.  Sorts last: first hardware
run of crc_pieces_kernel (written after the round's GPU budget was spent; its source is checked through the host emulation in
tests/test_crc.py)."""
import ctypes
import zlib

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _oracle():
    O = H.oracle()
    O.b2zo_crc64.restype = ctypes.c_uint64; O.b2zo_crc64.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
    return O


def test_gpu_digests_equal_the_oracle(pkg, codec):
    O = _oracle(); L = codec.L
    L.b200z_crc32_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint32)]
    L.b200z_crc64_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint64)]
    for name, data in list(H.sample_inputs(pkg, big=True).items()) + [("g2_64m", pkg.corpus.g2(64 << 20).tobytes())]:
        src = np.frombuffer(data, dtype=np.uint8) if data else np.zeros(1, dtype=np.uint8)
        c32, c64 = ctypes.c_uint32(1), ctypes.c_uint64(1)
        assert L.b200z_crc32_host(codec.h, src.ctypes.data, len(data), ctypes.byref(c32)) == 0
        assert L.b200z_crc64_host(codec.h, src.ctypes.data, len(data), ctypes.byref(c64)) == 0
        assert c32.value == zlib.crc32(data), name
        if len(data) <= (16 << 20):
            assert c64.value == O.b2zo_crc64(data, len(data)), name
