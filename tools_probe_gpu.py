import sys, time, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import __graft_entry__ as ge
pkg = ge.load_package()
n = int(sys.argv[1]) << 20
t=time.time(); data = pkg.corpus.g2(n); print('gen s', time.time()-t, 'cores', os.cpu_count())
src = torch.from_numpy(data).cuda()
c = pkg.Codec(0)
dst = torch.empty(c.compress_bound(n) + (n >> 10) + (1<<20), dtype=torch.uint8, device="cuda")
for it in range(3):
    c.reset_stats(); torch.cuda.synchronize(); t=time.time()
    m = c.compress_device(src.data_ptr(), n, dst.data_ptr(), dst.numel())
    dt=time.time()-t
    print(f'iter {it}: {n/1e6/dt:.0f} MB/s wall; match {c.stat(1):.2f} ms entropy {c.stat(2):.2f} ms assemble {c.stat(3):.2f} ms; ratio {n/m:.4f}')
for fl in (21, 20, 19, 18):
    c2 = pkg.Codec(0, frame_log=fl)
    for it in range(2):
        c2.reset_stats(); m = c2.compress_device(src.data_ptr(), n, dst.data_ptr(), dst.numel())
    print(f'frameLog {fl}: match {c2.stat(1):.2f} ms entropy {c2.stat(2):.2f} ms assemble {c2.stat(3):.2f}; ratio {n/m:.4f}')
    c2.close()
