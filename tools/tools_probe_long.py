"""cfg3 probe: long mode on G3 data (text + far copies), device-resident timing + ratio.  args: MiB windowLog"""
import sys, time, importlib
import numpy as np, torch
sys.path.insert(0, '.')
pkg = importlib.import_module('7-zip-zstd_b200')
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
wl = int(sys.argv[2]) if len(sys.argv) > 2 else 27
n = mib << 20
t = time.time(); host = pkg.corpus.g3(n); print(f"g3 {mib} MiB in {time.time()-t:.1f}s", flush=True)
src = torch.from_numpy(host).cuda()
for mode in (0, wl):
    c = pkg.Codec(0, long=mode) if mode else pkg.Codec(0)
    dst = torch.empty(c.compress_bound(n), dtype=torch.uint8, device='cuda')
    out = torch.empty(n + 64, dtype=torch.uint8, device='cuda')
    for rep in range(2):
        c.reset_stats(); torch.cuda.synchronize(); t = time.time()
        m = c.compress_device(src.data_ptr(), n, dst.data_ptr(), dst.numel()); torch.cuda.synchronize(); te = time.time() - t
        st = {k: round(c.stat(i), 1) for k, i in (("find", 1), ("dp", 10), ("entropy", 2), ("assemble", 3), ("dec_ent", 4), ("dec_exec", 5))}
        t = time.time(); r = c.decompress_device(dst.data_ptr(), m, out.data_ptr(), n); torch.cuda.synchronize(); td = time.time() - t
    ok = bool(torch.equal(out[:n], src))
    print(f"long={mode}: ratio {n/m:.4f} comp {m} enc {te*1e3:.1f} ms ({n/te/1e9:.1f} GB/s) dec {td*1e3:.1f} ms ({n/td/1e9:.1f} GB/s) ok {ok} stats {st}", flush=True)
    c.close()
