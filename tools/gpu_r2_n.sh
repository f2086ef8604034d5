#!/bin/bash
# tcgen05 kernel after the smem-traffic cut (row dot in the A converters, no sampler warps): regression + timing
mkdir -p gpurun_out
timeout 900 python tools/t32_check.py > gpurun_out/n_check.log 2>&1; tail -60 gpurun_out/n_check.log | cut -c1-260
CTN_TF32_NO_TMA=1 timeout 600 python tools/t32_check.py 0 1 2 3 4 5 6 7 8 9 10 11 12 13 14 16 17 18 19 > gpurun_out/n_check_notma.log 2>&1; grep -c "rel err" gpurun_out/n_check_notma.log; grep -A0 "OUT OF\|done,\|TFLOP" gpurun_out/n_check_notma.log | cut -c1-200
CTN_TF32_NO_TMA_A=1 timeout 600 python tools/t32_check.py 0 1 2 8 10 11 12 21 > gpurun_out/n_check_notmaa.log 2>&1; grep -A0 "OUT OF\|done,\|TFLOP" gpurun_out/n_check_notmaa.log | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 50 --warmup 5 --configs cfg3 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1: value', r['value'], 'ms/step', r['ms_per_step'])
for c in r.get('configs', []): print(json.dumps(c)[:1500])"
CTN_T32_DEBUG=16 timeout 300 python tools/t32_trace.py 21 > gpurun_out/n_trace21.log 2>&1; tail -8 gpurun_out/n_trace21.log
