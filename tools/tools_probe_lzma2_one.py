"""one LZMA2 decode launch for ncu (stream made by the reference encoder)"""
import os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [_ROOT, os.path.join(_ROOT, "tests")]
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
import helpers as H
import __graft_entry__ as ge
import torch
pkg = ge.load_package(); c = pkg.Codec(0)
mib = int(sys.argv[1]); mode = int(sys.argv[2])
raw = pkg.corpus.g2(mib << 20).tobytes()
prop, comp = H.ref_lzma2_compress(raw, 1, dict_size=1 << 20, block_size=1 << 20, threads=64)
size, nblk, used = c.lzma2_stream_info(comp)
d_src = torch.frombuffer(bytearray(comp) + bytearray(64), dtype=torch.uint8).cuda()
d_dst = torch.empty(size + 64, dtype=torch.uint8, device="cuda")
c.set("lzma2_model", mode)
c.lzma2_decompress_device(d_src.data_ptr(), len(comp), prop, d_dst.data_ptr(), size)
print("done", nblk)
