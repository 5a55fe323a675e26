#!/bin/bash
mkdir -p gpurun_out
B200Z_TRACE=1 timeout 600 python tools/tools_probe_e2e.py 4096 > gpurun_out/c10_e2e.txt 2>&1
grep -v "^\[b200z" gpurun_out/c10_e2e.txt | tail -8
grep "^\[b200z" gpurun_out/c10_e2e.txt | tail -12
