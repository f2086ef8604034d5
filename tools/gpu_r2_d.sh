#!/bin/bash
# round 2, GPU call D: sampler warps + by-value descriptors; fixed-cost analysis
mkdir -p gpurun_out
L=gpurun_out/d_t32.log
echo "== v2 ==" > $L
for i in $(seq 0 23); do timeout 150 python tools/t32_check.py $i >> $L 2>&1 || echo "case $i: exit $?" >> $L; done
echo "== v2 without TMA ==" >> $L
for i in 10 12 21; do CTN_TF32_NO_TMA=1 timeout 150 python tools/t32_check.py $i >> $L 2>&1 || echo "case $i: exit $?" >> $L; done
echo "== timing experiments (wrong results on purpose): no sampling (4), no lo conversion (8), both (12) ==" >> $L
for d in 4 8 12; do for i in 10 21 23; do echo "debug $d" >> $L; CTN_T32_DEBUG=$d timeout 150 python tools/t32_check.py $i >> $L 2>&1; done; done
grep -v "^ *$" $L | grep -v "^done" | tail -75
echo "== configs =="
timeout 900 python benchmarks/configs.py cfg3s cfg3 > gpurun_out/d_cfg3.log 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/d_cfg3.log'):
    try: r=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(r['config'], 'ms', round(r['ms'],3), 'TF', round(r['path_tflops'],1), 'by kind', {k: round(v,3) for k,v in r['ms_by_kind'].items()}, 'errT', r.get('max_rel_err_T'), 'errc', r.get('abs_err_log_scale'))
PY
echo "== ncu: tiny (22), one tile per CTA (23), two fused sites (21) =="
for i in 22 23 21; do
timeout 150 python tools/t32_check.py $i > gpurun_out/d_plain$i.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 4 -c 2 -o gpurun_out/d_t32_case$i python tools/t32_check.py $i > gpurun_out/d_ncu$i.log 2>&1; tail -1 gpurun_out/d_ncu$i.log
done
