#!/bin/bash
mkdir -p gpurun_out
echo "== region = frame"; timeout 600 python tools/tools_probe_e2e.py 4096 2>&1 | tail -4
echo "== region 19"; B200Z_REGION=19 timeout 600 python tools/tools_probe_e2e.py 4096 2>&1 | tail -4
echo "== region 18"; B200Z_REGION=18 timeout 600 python tools/tools_probe_e2e.py 4096 2>&1 | tail -4
