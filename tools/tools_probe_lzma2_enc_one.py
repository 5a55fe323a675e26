"""one LZMA2 encode (stage M + stage R) for ncu"""
import os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [_ROOT, os.path.join(_ROOT, "tests")]
import os, sys
import __graft_entry__ as ge
import torch
pkg = ge.load_package(); c = pkg.Codec(0)
mib = int(sys.argv[1]); c.set("lzma2_model", int(sys.argv[2]))
data = pkg.corpus.g2(mib << 20); n = data.nbytes
d_src = torch.from_numpy(data).cuda(); cap = c.lzma2_compress_bound(n)
d_dst = torch.empty(cap, dtype=torch.uint8, device="cuda")
sz, prop = c.lzma2_compress_device(d_src.data_ptr(), n, d_dst.data_ptr(), cap)
print("done", sz)
