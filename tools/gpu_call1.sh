#!/bin/bash
# first GPU call of round 2: stage F / stage G parity on hardware, timings, one full ncu capture of each
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zstd_enc.py -x -q > gpurun_out/c1_tests.txt 2>&1; echo "tests exit $?" >> gpurun_out/c1_tests.txt
tail -15 gpurun_out/c1_tests.txt
timeout 300 python tools/tools_profile_enc.py 4096 20 3 > gpurun_out/c1_stages.txt 2>&1; cat gpurun_out/c1_stages.txt
for cl in 6 8; do echo chunkLog $cl; timeout 300 python tools/tools_profile_enc.py 4096 20 2 $cl 2>&1 | tail -1; done
timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; tail -c 3000 gpurun_out/c1_bench.json; tail -5 gpurun_out/c1_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'zstd_enc_find_kernel|zstd_enc_dp_kernel' -c 2 -f -o gpurun_out/r2a_find_dp python tools/tools_profile_enc.py 1024 20 1 > gpurun_out/c1_ncu.txt 2>&1; tail -5 gpurun_out/c1_ncu.txt
ls -la gpurun_out | tail
