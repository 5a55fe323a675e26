#!/bin/bash
# whole GPU suite + bench + full ncu captures of stage F, G, E
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c2_tests.txt 2>&1; echo "tests exit $?" >> gpurun_out/c2_tests.txt
tail -12 gpurun_out/c2_tests.txt
timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; tail -c 2500 gpurun_out/c2_bench.json; tail -5 gpurun_out/c2_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'zstd_enc_find_kernel|zstd_enc_dp_kernel|zstd_enc_entropy_kernel' -c 3 -f -o gpurun_out/r2b_find_dp_ent python tools/tools_profile_enc.py 1024 20 1 > gpurun_out/c2_ncu.txt 2>&1; tail -3 gpurun_out/c2_ncu.txt
