#!/bin/bash
mkdir -p gpurun_out
for v in "" 1; do
echo "== CTN_TF32_NO_PDL=$v"
CTN_TF32_NO_PDL=$v timeout 600 python bench.py --steps 50 --warmup 5 --configs cfg3 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1: value', r['value'], 'ms/step', r['ms_per_step'], r['clocks'])
for c in r.get('configs', []): print(c['ms_per_contraction'], c['roofline']['frac'], c['errors'], c['ms_by_kernel_class'])"
done
timeout 600 python tools/t32_check.py 0 2 8 12 15 21 2>&1 | grep "OUT OF\|done,\|TFLOP\|flags [1-9]"
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -3
