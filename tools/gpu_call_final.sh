#!/bin/bash
# round-2 final evidence: smoke, the whole GPU suite, the default bench line, the launch list of the bench command
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; tail -c 300 gpurun_out/r2c_bench_n1.json; tail -2 gpurun_out/r2c_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-files-extra --no-lzma2-extra --no-refstreams-extra --no-long-extra > gpurun_out/r2c_launches_bench.log 2>&1
tail -c 200 gpurun_out/r2c_launches_bench.log
