#!/bin/bash
# lock-step stage R on hardware: parity tests of the method-21 encoder (all model placements), stage J's test again, timings at 4 GiB
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_zstd_dec.py tests/test_gpu_lzma2_enc.py -x -q 2>&1 | tail -8
timeout 600 python tools/tools_probe_lzma2_enc.py 4096 0,2 2>&1 | tail -4
