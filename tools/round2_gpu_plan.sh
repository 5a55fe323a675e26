#!/bin/bash
# First GPU call of the next round, in the order that makes every later minute count.  Run from the repo root under gpurun:
#   gpurun --timeout 1500 -- 'bash tools/round2_gpu_plan.sh > gpurun_out/round2_plan.log 2>&1; tail -40 gpurun_out/round2_plan.log'
# 1. hardware run of the kernels that have only seen the host emulator (stage Z), then the whole GPU suite
# 2. 4 GiB bench lines of the three parse modes that have none yet
# 3. launch lists + one full ncu capture each of stage P and stage Z (source-level hotspots decide what to optimise)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_lzma2_parse.py tests/test_gpu_zz_zstd_parse.py tests/test_gpu_zzz_crc.py tests/test_gpu_zzz_filters.py tests/test_gpu_zzz_xz.py -q
timeout 900 python -m pytest tests -m gpu -x -q
timeout 300 python tools/tools_probe_lzma2_parse.py 4096 20 2
timeout 300 python tools/tools_probe_lzma2_parse.py 4096 23 3
timeout 400 python bench.py --codec lzma2 --lzma2-parse 1 --steps 2 --warmup 1 --no-e2e > gpurun_out/bench_lzma2_parse1.json
timeout 400 python bench.py --level 12 --steps 2 --warmup 1 --no-e2e > gpurun_out/bench_zstd_level12.json
for c in lzma2 zstd; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_parse_$c.csv python tools/tools_profile_parse.py $c 1024
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lzma2_parse_kernel -c 1 -o gpurun_out/stage_p_full python tools/tools_profile_parse.py lzma2 256
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zstd_enc_parse_kernel -c 1 -o gpurun_out/stage_z_full python tools/tools_profile_parse.py zstd 256
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lzma2_cand_kernel -c 1 -o gpurun_out/stage_c_full python tools/tools_profile_parse.py lzma2 256
