#!/bin/bash
# 2 GPUs: peer-mailbox merge correctness + weak-scaling bench line
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 benchmarks/check_batch_shards.py > gpurun_out/k_check.log 2>&1; tail -4 gpurun_out/k_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/k_n2.json 2> gpurun_out/k_n2.err; tail -3 gpurun_out/k_n2.err
python - <<'PY'
import json
for f in ('gpurun_out/k_n2.json',):
    try:
        r=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'n', r['n_gpus'], 'value', r['value'], 'ms/step', r['ms_per_step'], 'e2e', r['e2e']['value'])
    except Exception as e: print(f, 'parse failed', e)
PY
CTN_NO_PEER_MERGE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 100 --warmup 5 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('NCCL merge: value', r['value'], 'ms/step', r['ms_per_step'])"
timeout 300 python bench.py --steps 100 --warmup 5 --configs none 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1: value', r['value'], 'ms/step', r['ms_per_step'])"
