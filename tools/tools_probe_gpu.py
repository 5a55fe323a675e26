"""Experiment driver: stage timings for different table sizes / frame sizes (device resident)."""
import os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [_ROOT, os.path.join(_ROOT, "tests")]
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
n = int(sys.argv[1]) << 20
src = torch.from_numpy(pkg.corpus.g2(n)).cuda()
dst = torch.empty(n + (n >> 6) + (1 << 20), dtype=torch.uint8, device="cuda")
back = torch.empty(n, dtype=torch.uint8, device="cuda")
for spec in sys.argv[2:]:
    fl, hl, hs = (int(x) for x in spec.split(","))
    c = pkg.Codec(0, frame_log=fl, hash_log_l=hl, hash_log_s=hs)
    for it in range(2):
        c.reset_stats(); m = c.compress_device(src.data_ptr(), n, dst.data_ptr(), dst.numel())
    k = c.decompress_device(dst.data_ptr(), m, back.data_ptr(), n)
    print(f"fl={fl} hl={hl} hs={hs}: match {c.stat(1):.1f} ms entropy {c.stat(2):.1f} ms | dec prepass {c.stat(9):.1f} entropy {c.stat(4):.1f} exec {c.stat(5):.1f} ms | ratio {n/m:.4f} ok={k==n}")
    c.close()
