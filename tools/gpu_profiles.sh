#!/bin/bash
# round-2 evidence: launch list of the bench command, DRAM traffic of the encoder kernels at the bench geometry (4 GiB), LZMA2 line
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_launches_bench.log 2>&1
tail -2 gpurun_out/r2_launches_bench.log | cut -c1-300
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'zstd_enc_find_kernel|zstd_enc_dp_kernel|zstd_enc_entropy_kernel|zstd_dec_' -c 12 --csv --log-file gpurun_out/r2_traffic_4g.csv python tools/tools_profile_enc.py 4096 20 1 > gpurun_out/r2_traffic_4g.log 2>&1
tail -2 gpurun_out/r2_traffic_4g.log
timeout 900 python bench.py --codec lzma2 --lzma2-parse 1 --frame-log 23 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_lzma2_parse1_f23.json 2> gpurun_out/r2_lzma2_parse1_f23.err; tail -c 1500 gpurun_out/r2_lzma2_parse1_f23.json; tail -3 gpurun_out/r2_lzma2_parse1_f23.err
timeout 600 python bench.py --codec lzma2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_lzma2_parse0.json 2> gpurun_out/r2_lzma2_parse0.err; tail -c 1200 gpurun_out/r2_lzma2_parse0.json
