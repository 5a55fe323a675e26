"""GPU probe: is the zstd decode stage time stable across repeated calls / alternation with encode?"""
import os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [_ROOT, os.path.join(_ROOT, "tests")]
import sys, torch
import __graft_entry__ as ge
pkg = ge.load_package()
n = int(sys.argv[1]) << 20
src = torch.from_numpy(pkg.corpus.g2(n)).cuda()
c = pkg.Codec(0)
cap = c.compress_bound(n)
dst = torch.empty(cap, dtype=torch.uint8, device="cuda"); back = torch.empty(n, dtype=torch.uint8, device="cuda")
m = c.compress_device(src.data_ptr(), n, dst.data_ptr(), cap)
def dec(tag):
    c.reset_stats(); k = c.decompress_device(dst.data_ptr(), m, back.data_ptr(), n)
    print(f"{tag}: prepass {c.stat(9):.1f} entropy {c.stat(4):.1f} exec {c.stat(5):.1f} ms", flush=True)
for i in range(4): dec(f"dec#{i}")
for i in range(3):
    c.compress_device(src.data_ptr(), n, dst.data_ptr(), cap); dec(f"after enc#{i}")
torch.cuda.empty_cache()
x = torch.empty(8 << 30, dtype=torch.uint8, device="cuda"); x.fill_(1); del x
for i in range(2): dec(f"after 8GiB fill#{i}")
