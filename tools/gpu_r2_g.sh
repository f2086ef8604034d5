#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/g_t32.log
echo "== v2 (TMA A+B) ==" > $L
for i in $(seq 0 23); do timeout 150 python tools/t32_check.py $i >> $L 2>&1 || echo "case $i: exit $?" >> $L; done
echo "== v2 without TMA ==" >> $L
for i in 4 12 21; do CTN_TF32_NO_TMA=1 timeout 150 python tools/t32_check.py $i >> $L 2>&1 || echo "case $i: exit $?" >> $L; done
grep -v "^ *$" $L | grep -v "^done" | cut -c1-220 | grep -E "kernel|TOLER|exit|==|err [0-9.e+-]+ " | tail -60
for i in 22 21; do timeout 120 python tools/t32_trace.py $i > gpurun_out/g_trace$i.log 2>&1; done
tail -14 gpurun_out/g_trace21.log
echo "== configs =="
timeout 900 python benchmarks/configs.py cfg3s cfg3 > gpurun_out/g_cfg3.log 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/g_cfg3.log'):
    try: r=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(r['config'], 'ms', round(r['ms'],3), 'TF', round(r['path_tflops'],1), 'by kind', {k: round(v,3) for k,v in r['ms_by_kind'].items()}, 'errT', r.get('max_rel_err_T'), 'errc', r.get('abs_err_log_scale'))
PY
timeout 1200 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_interp.py -q -m gpu 2>&1 | tail -12
