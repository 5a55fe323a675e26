#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_zstd_enc.py tests/test_ref_7z_host.py tests/test_boundary.py tests/test_gpu_zstd_dec.py tests/test_gpu_lzma2_dec.py -m gpu -x -q > gpurun_out/c3_tests.txt 2>&1; echo "tests exit $?" >> gpurun_out/c3_tests.txt
tail -25 gpurun_out/c3_tests.txt
timeout 300 python tools/tools_probe_e2e.py 4096 2>&1 | tail -8
