#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/c12_tests.txt 2>&1; echo "tests exit $?" >> gpurun_out/c12_tests.txt
tail -6 gpurun_out/c12_tests.txt
( time timeout 1200 python bench.py > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err ) 2>&1 | tail -3
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/c12_bench.json').read().strip().splitlines()[-1])
    print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'cpu', round(d['cpu_baseline']['value']), 'frac', round(d['roofline']['frac'], 4), 'traffic', d['roofline']['traffic'])
    print('kernel ms', d['config']['kernel_ms_per_step'])
    x = d['extra']
    print('lzma2', {k: (round(v, 1) if isinstance(v, float) else v) for k, v in x['lzma2'].items() if k not in ('workload', 'cpu_baseline', 'kernel_ms')})
    print('files', {k: (round(v, 1) if isinstance(v, float) else v) for k, v in x['many_files_7z'].items() if k != 'workload'})
    print('long', json.dumps(x.get('long_range'))[:1500])
    print('refstreams', json.dumps(x.get('reference_streams'))[:900])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/c12_bench.err').read()[-1500:])
PY
