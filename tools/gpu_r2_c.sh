#!/bin/bash
# round 2, GPU call C: loader/converter split
mkdir -p gpurun_out
L=gpurun_out/c_t32.log
echo "== v2 ==" > $L
for i in $(seq 0 23); do timeout 150 python tools/t32_check.py $i >> $L 2>&1 || echo "case $i: exit $?" >> $L; done
echo "== v2 without TMA ==" >> $L
for i in 10 11 12 21; do CTN_TF32_NO_TMA=1 timeout 150 python tools/t32_check.py $i >> $L 2>&1 || echo "case $i: exit $?" >> $L; done
echo "== v2 timing experiments: no sampling (4), no lo conversion (8), both (12): results are wrong on purpose ==" >> $L
for d in 4 8 12; do for i in 10 21; do echo "debug $d" >> $L; CTN_T32_DEBUG=$d timeout 150 python tools/t32_check.py $i >> $L 2>&1; done; done
echo "== v1 ==" >> $L
for i in 21 22 23; do CTN_TF32_V1=1 timeout 150 python tools/t32_check.py $i >> $L 2>&1; done
grep -v "^ *$" $L | grep -v "^done" | tail -75
echo "== configs =="
timeout 900 python benchmarks/configs.py cfg3s cfg3 > gpurun_out/c_cfg3.log 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/c_cfg3.log'):
    try: r=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(r['config'], 'ms', round(r['ms'],3), 'TF', round(r['path_tflops'],1), 'by kind', {k: round(v,3) for k,v in r['ms_by_kind'].items()}, 'errT', r.get('max_rel_err_T'), 'errc', r.get('abs_err_log_scale'))
PY
echo "== ncu 4096^3 =="
timeout 150 python tools/t32_check.py 10 > gpurun_out/c_plain10.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 3 -c 1 -o gpurun_out/c_t32_4096 python tools/t32_check.py 10 > gpurun_out/c_ncu10.log 2>&1; tail -2 gpurun_out/c_ncu10.log
