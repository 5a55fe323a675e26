#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_zstd_dec.py tests/test_gpu_zstd_enc.py tests/test_ref_7z_host.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/tools_probe_e2e.py 4096 2>&1 | tail -6
timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    print('value', round(d['value']), 'enc', round(d['config']['enc_MBps']), 'dec', round(d['config']['dec_MBps']), d['config']['kernel_ms_per_step'])
"
