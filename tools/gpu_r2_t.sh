#!/bin/bash
# N = 4: weak-scaling bench line + the new parity tests
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "default_optimizer" 2>&1 | tail -2
bash tools/gpu_r2_m.sh 4
