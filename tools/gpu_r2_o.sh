#!/bin/bash
# tcgen05 kernel with dedicated TMA-issuer warps + deferred row dot: regression, timing, load-skipping experiments
mkdir -p gpurun_out
timeout 900 python tools/t32_check.py > gpurun_out/o_check.log 2>&1; grep "OUT OF\|done,\|TFLOP\|flags [1-9]" gpurun_out/o_check.log | cut -c1-200
CTN_TF32_NO_TMA=1 timeout 600 python tools/t32_check.py 0 1 2 3 4 5 6 7 9 13 14 16 17 18 19 > gpurun_out/o_check_notma.log 2>&1; grep "OUT OF\|done,\|flags [1-9]" gpurun_out/o_check_notma.log | cut -c1-200
CTN_TF32_NO_TMA_A=1 timeout 600 python tools/t32_check.py 0 1 2 8 12 > gpurun_out/o_check_notmaa.log 2>&1; grep "OUT OF\|done,\|flags [1-9]" gpurun_out/o_check_notmaa.log | cut -c1-200
for f in 64 128 192 196 204; do echo "== CTN_T32_DEBUG=$f (64 no A loads, 128 no B loads, 4 no row sums, 8 no lo conversion)"; CTN_T32_DEBUG=$f timeout 300 python tools/t32_check.py 10 11 12 2>&1 | grep TFLOP; done
CTN_T32_DEBUG=16 timeout 300 python tools/t32_trace.py 21 > gpurun_out/o_trace21.log 2>&1; tail -4 gpurun_out/o_trace21.log
CTN_T32_DEBUG=16 timeout 300 python tools/t32_trace.py 10 > gpurun_out/o_trace10.log 2>&1; tail -2 gpurun_out/o_trace10.log
timeout 600 python bench.py --steps 50 --warmup 5 --configs cfg3 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1: value', r['value'], 'ms/step', r['ms_per_step'])
for c in r.get('configs', []): print(c['ms_per_contraction'], c['roofline']['frac'], c['errors'])"
