#!/bin/bash
# the bench as the driver launches it at N = 2 (NUMA binding of the ranks, one_call_multi_gpu extra)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err
tail -c 1500 gpurun_out/r2b_bench_n2.json; tail -5 gpurun_out/r2b_bench_n2.err
