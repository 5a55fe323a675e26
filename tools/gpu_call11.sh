#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zstd_dec.py tests/test_gpu_zstd_long.py tests/test_gpu_zstd_enc.py tests/test_ref_7z_host.py -q > gpurun_out/c11_tests.txt 2>&1; echo "tests exit $?" >> gpurun_out/c11_tests.txt
tail -5 gpurun_out/c11_tests.txt
timeout 600 python tools/tools_probe_long.py 4096 27 > gpurun_out/c11_long.txt 2>&1; tail -3 gpurun_out/c11_long.txt
timeout 600 python tools/tools_probe_e2e.py 4096 > gpurun_out/c11_e2e.txt 2>&1; tail -5 gpurun_out/c11_e2e.txt
