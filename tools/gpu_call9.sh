#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zstd_long.py -q > gpurun_out/c9_tests.txt 2>&1; echo "tests exit $?" >> gpurun_out/c9_tests.txt
tail -8 gpurun_out/c9_tests.txt
timeout 900 python tools/tools_probe_long.py 8192 27 > gpurun_out/c9_long.txt 2>&1; tail -5 gpurun_out/c9_long.txt
