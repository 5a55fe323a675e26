#!/bin/bash
# stage J on hardware: decoder parity tests, then a reference-written 1 GiB single frame through both execute paths, then the launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_zstd_dec.py -x -q 2>&1 | tail -8
timeout 400 python tools/tools_probe_jump.py 1024 1,0 2>&1 | tail -8
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jump --csv --log-file gpurun_out/launches_jump.csv python tools/tools_probe_jump.py 256 1 > gpurun_out/probe_jump_ncu.log 2>&1
tail -3 gpurun_out/probe_jump_ncu.log
