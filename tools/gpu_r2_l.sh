#!/bin/bash
# N = 2: bench line with the multi-GPU sub-records
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/l_n2.json 2> gpurun_out/l_n2.err; tail -3 gpurun_out/l_n2.err
python - <<'PY'
import json
r=json.loads(open('gpurun_out/l_n2.json').read().strip().splitlines()[-1]); print('n', r['n_gpus'], 'value', r['value'], 'ms/step', r['ms_per_step'], 'e2e', r['e2e']['value'])
print(json.dumps(r.get('multi_gpu'), indent=1)[:2500])
PY
