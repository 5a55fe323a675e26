"""GPU probe: cfg5 shape -- 100k files of 64 KiB, mixed entropy, one batch call"""
import os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [_ROOT, os.path.join(_ROOT, "tests")]
import sys, time
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); c = pkg.Codec(0)
nfiles = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
fs = 65536
text = pkg.corpus.g2(nfiles * fs // 2)
parts = [text, pkg.corpus.entropy_class(1, nfiles * fs // 8), pkg.corpus.entropy_class(2, nfiles * fs // 8), pkg.corpus.entropy_class(3, nfiles * fs // 4)]
buf = np.concatenate(parts)[: nfiles * fs]
import ctypes, torch
sizes = np.full(nfiles, fs, dtype=np.uint64)
cap = c.L.b200z_zstd_compress_batch_bound(c.h, buf.nbytes, nfiles)
hin = torch.from_numpy(buf).pin_memory(); hout = torch.empty(cap, dtype=torch.uint8).pin_memory(); offs = np.zeros(nfiles + 1, dtype=np.uint64)
for it in range(2):
    t = time.time()
    rc = c.L.b200z_zstd_compress_batch_host(c.h, hin.data_ptr(), sizes.ctypes.data, nfiles, hout.data_ptr(), cap, offs.ctypes.data)
    t = time.time() - t
    assert rc == 0
back = torch.empty(buf.nbytes, dtype=torch.uint8).pin_memory()
td = time.time(); n = c.decompress_into(hout.data_ptr(), int(offs[-1]), back.data_ptr(), buf.nbytes); td = time.time() - td
print(f"files={nfiles} x {fs} B: batch compress {t*1e3:.0f} ms -> {buf.nbytes/t/1e6:.0f} MB/s ({nfiles/t:.0f} files/s), ratio {buf.nbytes/int(offs[-1]):.3f}; "
      f"decompress whole {td*1e3:.0f} ms -> {buf.nbytes/td/1e6:.0f} MB/s ok={n == buf.nbytes and bool(torch.equal(back, hin))}")
