"""GPU driver for profiling the price-based parses (run under `ncu ... python tools/tools_profile_parse.py <codec> <MiB> [frameLog] [sliceLog]`):
one encode of G2 text through the device-pointer entry, so the launch list holds stage C, stage P / stage Z, stage R / stage E."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
pkg = g.load_package()
codec = sys.argv[1] if len(sys.argv) > 1 else "lzma2"
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
fl = int(sys.argv[3]) if len(sys.argv) > 3 else 20
sl = int(sys.argv[4]) if len(sys.argv) > 4 else 2
data = pkg.corpus.g2(mb << 20)
d_src = torch.from_numpy(data).cuda()
if codec == "lzma2":
    c = pkg.Codec(0, frame_log=fl, window_log=fl, lzma2_slice_log=sl, lzma2_parse=1)
    cap = c.lzma2_compress_bound(data.nbytes); d_dst = torch.empty(cap, dtype=torch.uint8, device="cuda")
    n, prop = c.lzma2_compress_device(d_src.data_ptr(), data.nbytes, d_dst.data_ptr(), cap)
else:
    c = pkg.Codec(0, frame_log=fl, window_log=fl, zstd_parse=1)
    cap = c.compress_bound(data.nbytes); d_dst = torch.empty(cap, dtype=torch.uint8, device="cuda")
    n = c.compress_device(d_src.data_ptr(), data.nbytes, d_dst.data_ptr(), cap)
print(f"{codec} price-based parse, {mb} MiB, frameLog {fl}: ratio {data.nbytes / n:.4f}; stage C {c.stat(1):.0f} ms, parse {c.stat(10):.0f} ms, "
      f"coder {c.stat(2):.0f} ms, assemble {c.stat(3):.0f} ms")
