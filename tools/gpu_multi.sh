#!/bin/bash
# multi-GPU checks of the dispatcher inside the product (run with gpurun --gpus N)
N=${1:-2}
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_zstd_enc.py -m gpu -x -q -k "device_count" 2>&1 | tail -3
timeout 600 python bench.py --one-process --gpus $N --steps 3 2>&1 | tail -1 | cut -c1-900
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:400]); continue
    print('N', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'extra', json.dumps(d.get('extra'))[:700])
"
