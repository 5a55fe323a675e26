#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_zstd_enc.py tests/test_gpu_lzma2_enc.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/tools_profile_enc.py 4096 20 3 2>&1 | tail -2
python - <<'PY'
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch, __graft_entry__ as ge
pkg = ge.load_package()
n = 4096 << 20
src = torch.from_numpy(pkg.corpus.g2(n)).cuda()
for level in (1, 3, 6):
    c = pkg.Codec(0, level=level)
    dst = torch.empty(c.compress_bound(n), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        c.reset_stats(); m = c.compress_device(src.data_ptr(), n, dst.data_ptr(), dst.numel())
    print(f"level {level}: find {c.stat(1):.2f} ms parse {c.stat(10):.2f} ms entropy {c.stat(2):.2f} ms ratio {n/m:.4f}")
    c.close()
PY
