"""Profiling driver: one compress of N MiB of G2 text, device resident (used under ncu)."""
import os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [_ROOT, os.path.join(_ROOT, "tests")]
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
n = int(sys.argv[1]) << 20
fl = int(sys.argv[2]) if len(sys.argv) > 2 else 20
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
src = torch.from_numpy(pkg.corpus.g2(n)).cuda()
cl = int(sys.argv[4]) if len(sys.argv) > 4 else 7
c = pkg.Codec(0, frame_log=fl, chunk_log=cl)
dst = torch.empty(c.compress_bound(n), dtype=torch.uint8, device="cuda")
for _ in range(reps):
    c.reset_stats()
    m = c.compress_device(src.data_ptr(), n, dst.data_ptr(), dst.numel())
    print(f"find {c.stat(1):.2f} ms parse {c.stat(10):.2f} ms entropy {c.stat(2):.2f} ms assemble {c.stat(3):.2f} ms ratio {n/m:.4f}")
