#!/bin/bash
# N GPUs (N = $1): bench line with the multi-GPU sub-records
N=$1
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/m_n$N.json 2> gpurun_out/m_n$N.err; tail -5 gpurun_out/m_n$N.err | cut -c1-300
python - <<PY
import json
r=json.loads(open('gpurun_out/m_n$N.json').read().strip().splitlines()[-1]); print('n', r['n_gpus'], 'value', r['value'], 'ms/step', r['ms_per_step'], 'e2e', r['e2e']['value'], r['e2e'].get('h2d_GBps_per_gpu'))
print(json.dumps(r.get('multi_gpu'), indent=1)[:2500])
PY
