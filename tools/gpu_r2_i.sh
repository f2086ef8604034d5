#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/i_t32.log
echo "== v2 per-row debias ==" > $L
for i in $(seq 0 23); do timeout 150 python tools/t32_check.py $i >> $L 2>&1 || echo "case $i: exit $?" >> $L; done
for i in 4 12 21; do CTN_TF32_NO_TMA=1 timeout 150 python tools/t32_check.py $i >> $L 2>&1 || echo "case $i: exit $?" >> $L; done
grep -v "^ *$" $L | grep -v "^done" | cut -c1-200 | grep -E "kernel|TOLER|exit|==|err " | sed -E 's/^([0-9]+) .*T-hat/\1 T-hat/' | tail -50
python bench.py --steps 30 --warmup 5 --configs cfg3 > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; tail -2 gpurun_out/i_bench.err
python - <<'PY'
import json
try:
    r=json.loads(open('gpurun_out/i_bench.json').read().strip().splitlines()[-1])
    print('value', r['value'], 'ms/step', r['ms_per_step'], 'e2e', r['e2e']['value'], 'roofline', r['roofline']['frac'])
    for c in r.get('configs', []):
        print(json.dumps({k: c.get(k) for k in ('name','error','ms_per_contraction','path_tflops','errors')})[:600])
        print(json.dumps(c.get('roofline'))[:400])
except Exception as e:
    print('bench parse failed', e)
PY
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/i_pytest.log
