#!/bin/bash
# evidence after stage J: the whole GPU suite, the default bench line, a full ncu capture of stage J's kernels
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 900 python bench.py > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; tail -c 600 gpurun_out/r2b_bench_n1.json; tail -3 gpurun_out/r2b_bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jump -c 3 -o gpurun_out/r2_jump_full python tools/tools_probe_jump.py 512 1 > gpurun_out/jump_ncu.log 2>&1
tail -2 gpurun_out/jump_ncu.log
