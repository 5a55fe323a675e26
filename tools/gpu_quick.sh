#!/bin/bash
# quick hardware check while iterating on stage F / stage G: parity tests of the zstd encoder, then stage timings at 4 GiB
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zstd_enc.py -x -q 2>&1 | tail -3
timeout 300 python tools/tools_profile_enc.py 4096 20 3 2>&1 | tail -3
