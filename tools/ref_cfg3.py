"""The reference's own encoder (oracle/_ref/libref_zstd.so) on a sample of the cfg3 input (G3: text + far copies): ratio and speed of
`zstd:x{level}:long=27` as ZstdEncoder.cpp sets it (level, enableLongDistanceMatching, windowLog 27, nbWorkers = cores).
args: MiB level [jobSizeMiB].  Writes one JSON line (kept under profiles/ for bench.py's extra.long_range)."""
import sys, time, json, os, importlib
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import helpers
pkg = importlib.import_module('7-zip-zstd_b200')
mib = int(sys.argv[1]); level = int(sys.argv[2]); job = int(sys.argv[3]) if len(sys.argv) > 3 else 0
n = mib << 20
data = pkg.corpus.g3(n)
Z = helpers.ref()
cores = os.cpu_count() or 1
c = Z.ZSTD_createCCtx()
for k, v in ((100, level), (160, 1), (101, 27), (400, min(cores, 200))) + (((402, job << 20),) if job else ()):
    r = Z.ZSTD_CCtx_setParameter(c, k, v); assert not Z.ZSTD_isError(r), (k, v)
out = np.zeros(Z.ZSTD_compressBound(n), dtype=np.uint8)
t = time.perf_counter(); r = Z.ZSTD_compress2(c, out.ctypes.data, out.size, data.ctypes.data, n); te = time.perf_counter() - t
assert not Z.ZSTD_isError(r)
back = np.zeros(n, dtype=np.uint8)
t = time.perf_counter(); d = Z.ZSTD_decompress(back.ctypes.data, n, out.ctypes.data, r); td = time.perf_counter() - t
assert d == n and np.array_equal(back, data)
print(json.dumps({"sample_MiB": mib, "level": level, "long": 27, "nbWorkers": min(cores, 200), "jobSize_MiB": job or "default", "ratio": n / r,
                  "enc_MBps": n / 1e6 / te, "dec_MBps": n / 1e6 / td, "t_enc_s": te, "t_dec_s": td, "cores": cores}))
