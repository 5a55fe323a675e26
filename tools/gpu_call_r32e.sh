#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --section SourceCounters --section WarpStateStats --section InstructionStats --section LaunchStats --clock-control none --import-source on -k regex:range32 -c 1 -o gpurun_out/r2_range32_v3 python tools/tools_probe_lzma2_enc.py 512 3 > gpurun_out/range32_v3.log 2>&1
tail -2 gpurun_out/range32_v3.log
