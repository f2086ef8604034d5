#!/bin/bash
mkdir -p gpurun_out
(time python bench.py --steps 50 --warmup 5 > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err) 2> gpurun_out/h_bench.time
tail -3 gpurun_out/h_bench.err; cat gpurun_out/h_bench.time | tail -4
python - <<'PY'
import json
try:
    r=json.loads(open('gpurun_out/h_bench.json').read().strip().splitlines()[-1])
    print('value', r['value'], 'e2e', r['e2e']['value'], 'roofline', r['roofline']['frac'], 'cpu', r.get('cpu_baseline',{}).get('value'))
    print('peaks', r.get('peaks'))
    for c in r.get('configs', []):
        print(json.dumps({k: c.get(k) for k in ('name','error','ms_per_contraction','path_tflops','roofline','errors','cpu_baseline','e2e')})[:900])
except Exception as e:
    print('bench parse failed', e)
PY
(time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/h_bench_ref.json 2>&1) 2>&1 | tail -3; head -c 600 gpurun_out/h_bench_ref.json; echo
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/h_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
