#!/bin/bash
# stage R, 32 chains per warp, third version (working set in shared memory, producer prefetch): parity + timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lzma2_enc.py -x -q 2>&1 | tail -3
timeout 600 python tools/tools_probe_lzma2_enc.py 4096 3 2>&1 | tail -1
timeout 600 python tools/tools_probe_lzma2_enc.py 1024 3 2>&1 | tail -1
