#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_zstd_enc.py -m gpu -x -q -k ladder 2>&1 | grep -E "^E" | tail -6
python - <<'PY'
import sys, ctypes, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import helpers as H, __graft_entry__ as g
pkg = g.load_package()
inputs = H.sample_inputs(pkg, big=False)
data = inputs["mixed"] + inputs["g2_1m"] + b"0123456789abcdef" * 5000
for level in (1, 3, 5):
    p = H.EncParams(); H.oracle().b2zo_enc_default_params(ctypes.byref(p), level)
    c = pkg.Codec(0, level=level)
    got = c.stage_f(data)
    want = H.oracle_candidates(data, flags=p.flags, hashLogS=p.hashLogS)
    bad = np.nonzero(got != want)[0]
    print("level", level, "stage F mismatches", len(bad), bad[:6], [hex(int(got[i])) for i in bad[:3]], [hex(int(want[i])) for i in bad[:3]])
    comp = c.compress(data); o = H.oracle_compress(data, flags=p.flags, hashLogS=p.hashLogS)
    print("   frames equal:", comp == o, len(comp), len(o))
    c.close()
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:zstd_dec -c 40 python -c "
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch, __graft_entry__ as ge
pkg = ge.load_package(); n = 1024 << 20
src = torch.from_numpy(pkg.corpus.g2(n)).cuda(); c = pkg.Codec(0)
dst = torch.empty(c.compress_bound(n), dtype=torch.uint8, device='cuda'); back = torch.empty(n, dtype=torch.uint8, device='cuda')
m = c.compress_device(src.data_ptr(), n, dst.data_ptr(), dst.numel())
c.decompress_device(dst.data_ptr(), m, back.data_ptr(), n)
" 2>&1 | grep -E "zstd_dec_[a-z_<>0-9]*|gpu__time" | paste - - | awk '{print $1, $(NF)}' | sort | uniq -c | head -20
