#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_zstd_enc.py -m gpu -x -q -k ladder 2>&1 | grep -E "assert|Error|sizes|level" | head -12
timeout 1500 python -m pytest tests/test_gpu_zstd_dec.py tests/test_gpu_zzz_xz.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --no-lzma2-extra --steps 3 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'enc', round(d['config']['enc_MBps']), 'dec', round(d['config']['dec_MBps']), d['config']['kernel_ms_per_step'])
"
timeout 300 python tools/tools_probe_e2e.py 4096 2>&1 | tail -5
