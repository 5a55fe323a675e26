#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zstd_dec.py -x -q -k "stage_j" 2>&1 | tail -3
timeout 300 python tools/tools_probe_jump.py 1024 1 3 0 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jump -c 36 --csv --log-file gpurun_out/launches_jump2.csv python tools/tools_probe_jump.py 256 1 > /dev/null 2>&1
