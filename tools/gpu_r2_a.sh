#!/bin/bash
# round 2, GPU call A: bring-up of the TS-mode tcgen05 kernel (each case in its own process)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
echo "== v2 (default) ==" > gpurun_out/a_t32.log
for i in $(seq 0 20); do timeout 150 python tools/t32_check.py $i >> gpurun_out/a_t32.log 2>&1 || echo "case $i: exit $?" >> gpurun_out/a_t32.log; done
echo "== v2 without TMA ==" >> gpurun_out/a_t32.log
for i in 2 3 10 11 12 13; do CTN_TF32_NO_TMA=1 timeout 150 python tools/t32_check.py $i >> gpurun_out/a_t32.log 2>&1 || echo "case $i: exit $?" >> gpurun_out/a_t32.log; done
echo "== v1 (round-1 kernel) ==" >> gpurun_out/a_t32.log
for i in 8 10 12 15 20; do CTN_TF32_V1=1 timeout 150 python tools/t32_check.py $i >> gpurun_out/a_t32.log 2>&1 || echo "case $i: exit $?" >> gpurun_out/a_t32.log; done
grep -v "^ *$" gpurun_out/a_t32.log | tail -80
echo "== configs =="
timeout 900 python benchmarks/configs.py cfg3s cfg3 > gpurun_out/a_cfg3.log 2>&1; tail -5 gpurun_out/a_cfg3.log
CTN_TF32_V1=1 timeout 600 python benchmarks/configs.py cfg3s cfg3 > gpurun_out/a_cfg3_v1.log 2>&1; tail -3 gpurun_out/a_cfg3_v1.log
echo "== pytest gemm/parity =="
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/a_pytest.log
