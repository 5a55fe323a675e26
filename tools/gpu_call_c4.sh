#!/bin/bash
# stage C with a warp per (frame, table): parity tests of both price-based parses, stage times at 4 GiB / 8 MiB frames and at 1 MiB frames
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_zz_lzma2_parse.py tests/test_gpu_zz_zstd_parse.py -x -q 2>&1 | tail -4
timeout 600 python tools/tools_probe_lzma2_parse.py 4096 23 2 2>&1 | head -2
timeout 300 python tools/tools_probe_lzma2_parse.py 2048 20 2 2>&1 | head -1
