"""GPU probe: LZMA2 encode (stage M + stage R) and decode of our own streams, device-resident."""
import os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [_ROOT, os.path.join(_ROOT, "tests")]
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
import __graft_entry__ as ge
import torch
pkg = ge.load_package(); c = pkg.Codec(0)
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
data = pkg.corpus.g2(mib << 20)
n = data.nbytes
d_src = torch.from_numpy(data).cuda()
cap = c.lzma2_compress_bound(n)
d_dst = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
for model in ([int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else (0, 2)):
    c.set("lzma2_model", model)
    for it in range(2):
        c.reset_stats(); torch.cuda.synchronize(); t0 = time.time()
        sz, prop = c.lzma2_compress_device(d_src.data_ptr(), n, d_dst.data_ptr(), cap)
        torch.cuda.synchronize(); t = time.time() - t0
        m, r, a = c.stat(1), c.stat(2), c.stat(3)
    c.reset_stats()
    got = c.lzma2_decompress_device(d_dst.data_ptr(), sz, prop, d_out.data_ptr(), n)
    dec = c.stat(4); pre = c.stat(9)
    ok = got == n and bool(torch.equal(d_out[:n], d_src))
    print(f"model={model} n={mib}MiB ratio={n/sz:.4f} enc: match={m:.1f}ms range={r:.1f}ms assemble={a:.1f}ms wall={t*1e3:.1f}ms -> {n/t/1e6:.0f} MB/s | dec: walk={pre:.1f}ms decode={dec:.1f}ms -> {n/(pre+dec)/1e3:.0f} MB/s ok={ok}", flush=True)
