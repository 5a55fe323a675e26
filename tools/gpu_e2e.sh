#!/bin/bash
timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'enc', round(d['config']['enc_MBps']), 'dec', round(d['config']['dec_MBps']), d['config']['kernel_ms_per_step'])
"
