#!/bin/bash
# lock-step stage R with forced convergence + warp XXH64 verify: parity tests, timings, one full ncu capture of the range kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_zstd_dec.py tests/test_gpu_lzma2_enc.py -x -q 2>&1 | tail -5
timeout 600 python tools/tools_probe_lzma2_enc.py 4096 0 2>&1 | tail -2
timeout 300 python tools/tools_probe_jump.py 1024 1 3 1 2>&1 | tail -4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:range32 -c 1 -o gpurun_out/r2_range32_full python tools/tools_probe_lzma2_enc.py 1024 0 > gpurun_out/range32_ncu.log 2>&1
tail -2 gpurun_out/range32_ncu.log
