#!/bin/bash
# segmented stage J + warp XXH64 with precomputed inputs: decoder tests (2.5 GiB frame included), checksummed 1 GiB frame
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_zstd_dec.py -x -q 2>&1 | tail -5
timeout 300 python tools/tools_probe_jump.py 1024 1 3 1 2>&1 | tail -3
timeout 300 python tools/tools_probe_jump.py 1024 1 3 0 2>&1 | tail -2
