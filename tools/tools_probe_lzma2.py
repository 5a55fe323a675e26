"""GPU probe: LZMA2 decode throughput vs. number of independent blocks (streams made by the reference encoder)."""
import os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [_ROOT, os.path.join(_ROOT, "tests")]
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
import helpers as H
import __graft_entry__ as ge
import torch

pkg = ge.load_package()
c = pkg.Codec(0)
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
data = pkg.corpus.g2(mib << 20)
raw = data.tobytes()
ncpu = os.cpu_count()
for blk_log in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else '20,22').split(',')]:
    t0 = time.time()
    prop, comp = H.ref_lzma2_compress(raw, 1, dict_size=1 << blk_log, block_size=1 << blk_log, threads=min(ncpu, 128))
    t_enc = time.time() - t0
    size, nblk, used = c.lzma2_stream_info(comp)
    d_src = torch.frombuffer(bytearray(comp) + bytearray(64), dtype=torch.uint8).cuda()
    d_dst = torch.empty(size + 64, dtype=torch.uint8, device="cuda")
    res = {}
    for mode in (1, 2):
        c.set("lzma2_model", mode)
        for it in range(2):
            d_dst.zero_()
            c.reset_stats()
            n = c.lzma2_decompress_device(d_src.data_ptr(), len(comp), prop, d_dst.data_ptr(), size)
            pre, dec = c.stat(9), c.stat(4)
        res[mode] = dec
        ok = bool(torch.equal(d_dst[:size].cpu(), torch.from_numpy(data)))
        print(f"  mode={mode} walk={pre:.2f}ms decode={dec:.2f}ms -> {size/dec/1e3:.0f} MB/s ok={ok}", flush=True)
    sub = 64 << 20
    sprop, scomp = H.ref_lzma2_compress(raw[:sub], 1, dict_size=1 << blk_log, block_size=1 << blk_log, threads=min(ncpu, 128))
    t0 = time.time(); H.ref_lzma2_decompress(scomp, sub, sprop); t_ref = (time.time() - t0) * size / sub
    print(f"blk=2^{blk_log} blocks={nblk} ratio={size/len(comp):.3f} ref_enc={size/t_enc/1e6:.0f}MB/s({min(ncpu,128)}thr) "
          f"gpu_walk={pre:.2f}ms gpu_decode={dec:.2f}ms -> {size/dec/1e3:.0f} MB/s ok={ok} ref_dec_1thr={size/t_ref/1e6:.0f} MB/s", flush=True)
