#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_zstd_dec.py tests/test_gpu_lzma2_enc.py -x -q 2>&1 | tail -5
timeout 600 python tools/tools_probe_lzma2_enc.py 4096 0 2>&1 | tail -2
timeout 600 python tools/tools_probe_lzma2_enc.py 1024 0 2>&1 | tail -2
