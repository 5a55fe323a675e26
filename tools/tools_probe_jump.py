"""GPU probe: a reference-written single-frame zstd stream (sliding window) through the decoder, stage J against the execution units.
usage: tools_probe_jump.py <MiB> [modes e.g. 1,0] [level] [checksum 0/1]"""
import os, sys, time
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [_ROOT, os.path.join(_ROOT, "tests")]
import torch
import __graft_entry__ as ge
import helpers
pkg = ge.load_package()
n = int(sys.argv[1]) << 20
modes = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1").split(",")]
level = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cks = int(sys.argv[4]) if len(sys.argv) > 4 else 0
data = pkg.corpus.g2(n)
t = time.perf_counter(); comp = helpers.ref_compress(data.tobytes(), level, cks, nbWorkers=min(os.cpu_count() or 1, 64)); t = time.perf_counter() - t
print(f"reference level {level} checksum {cks}: {n >> 20} MiB -> {len(comp)} bytes in {t:.1f} s", flush=True)
src = torch.frombuffer(bytearray(comp + bytes(64)), dtype=torch.uint8).cuda()
back = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
want = torch.from_numpy(data).cuda()
for mode in modes:
    c = pkg.Codec(0, dec_jump=mode)
    for rep in range(3 if mode else 1):
        back.zero_(); c.reset_stats(); torch.cuda.synchronize()
        t = time.perf_counter(); k = c.decompress_device(src.data_ptr(), len(comp), back.data_ptr(), n); torch.cuda.synchronize(); t = time.perf_counter() - t
        ok = k == n and bool(torch.equal(back[:n], want))
        print(f"mode {mode} #{rep}: {t * 1e3:.1f} ms ({n / t / 1e9:.2f} GB/s) prepass {c.stat(9):.1f} entropy {c.stat(4):.1f} exec {c.stat(5):.1f} ms, jump frames {int(c.stat(11))}, exact {ok}", flush=True)
    c.close()
