"""Experiment driver: host-pointer (e2e) timings and raw PCIe copy bandwidth."""
import os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [_ROOT, os.path.join(_ROOT, "tests")]
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
n = int(sys.argv[1]) << 20
hin = torch.empty(n, dtype=torch.uint8).pin_memory()
pkg.corpus.g2_into(hin.data_ptr(), n)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(hin, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"H2D pinned {n/1e9/dt:.1f} GB/s")
hback = torch.empty(n, dtype=torch.uint8).pin_memory()
for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); hback.copy_(d, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"D2H pinned {n/1e9/dt:.1f} GB/s")
for hb in (28, 29, 30, 31):
    c = pkg.Codec(0, host_batch_log=hb, **({'region_log': int(os.environ['B200Z_REGION'])} if os.environ.get('B200Z_REGION') else {}))
    bound = c.compress_bound(n)
    hcomp = torch.empty(bound, dtype=torch.uint8).pin_memory()
    for it in range(2):
        t = time.perf_counter(); m = c.compress_into(hin.data_ptr(), n, hcomp.data_ptr(), bound); te = time.perf_counter() - t
        t = time.perf_counter(); k = c.decompress_into(hcomp.data_ptr(), m, hback.data_ptr(), n); td = time.perf_counter() - t
    print(f"host_batch_log={hb}: enc {te*1e3:.0f} ms ({n/1e9/te:.2f} GB/s) dec {td*1e3:.0f} ms ({n/1e9/td:.2f} GB/s) ok={torch.equal(hback, hin)}")
    c.close()
