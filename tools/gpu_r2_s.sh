#!/bin/bash
# tcgen05 kernel: regression (all layouts / fallbacks), unit + full-size tests, cfg3 bench, time lines
mkdir -p gpurun_out
timeout 900 python tools/t32_check.py > gpurun_out/s_check.log 2>&1; grep "OUT OF\|done,\|TFLOP\|flags [1-9]" gpurun_out/s_check.log | cut -c1-200
CTN_TF32_NO_TMA=1 timeout 600 python tools/t32_check.py 0 1 2 3 4 5 6 7 9 13 14 16 17 18 19 > gpurun_out/s_check_notma.log 2>&1; grep "OUT OF\|done,\|flags [1-9]" gpurun_out/s_check_notma.log | cut -c1-200
CTN_TF32_NO_TMA_A=1 timeout 600 python tools/t32_check.py 0 1 2 8 12 > gpurun_out/s_check_notmaa.log 2>&1; grep "OUT OF\|done,\|flags [1-9]" gpurun_out/s_check_notmaa.log | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/t32_trace.py 21 > gpurun_out/s_trace21.log 2>&1; tail -9 gpurun_out/s_trace21.log
timeout 300 python tools/t32_trace.py 22 > gpurun_out/s_trace22.log 2>&1; tail -12 gpurun_out/s_trace22.log
timeout 600 python bench.py --steps 50 --warmup 5 --configs cfg3 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1: value', r['value'], 'ms/step', r['ms_per_step'])
for c in r.get('configs', []): print(c['ms_per_contraction'], c['roofline']['frac'], c['errors'])"
