#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/f_t32.log
echo "== v2 (TMA A+B) ==" > $L
for i in $(seq 0 23); do timeout 150 python tools/t32_check.py $i >> $L 2>&1 || echo "case $i: exit $?" >> $L; done
echo "== v2, TMA for B only ==" >> $L
for i in 4 10 11 12 21 23; do CTN_TF32_NO_TMA_A=1 timeout 150 python tools/t32_check.py $i >> $L 2>&1 || echo "case $i: exit $?" >> $L; done
echo "== v2 without TMA ==" >> $L
for i in 4 10 12 21; do CTN_TF32_NO_TMA=1 timeout 150 python tools/t32_check.py $i >> $L 2>&1 || echo "case $i: exit $?" >> $L; done
grep -v "^ *$" $L | grep -v "^done" | cut -c1-200 | tail -80
for i in 22 23 21; do timeout 120 python tools/t32_trace.py $i > gpurun_out/f_trace$i.log 2>&1; done
head -30 gpurun_out/f_trace23.log
echo "== configs =="
timeout 900 python benchmarks/configs.py cfg3s cfg3 > gpurun_out/f_cfg3.log 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/f_cfg3.log'):
    try: r=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(r['config'], 'ms', round(r['ms'],3), 'TF', round(r['path_tflops'],1), 'by kind', {k: round(v,3) for k,v in r['ms_by_kind'].items()}, 'errT', r.get('max_rel_err_T'), 'errc', r.get('abs_err_log_scale'))
PY
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -5
