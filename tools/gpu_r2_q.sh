#!/bin/bash
# A-converter time line + row-dot variants
mkdir -p gpurun_out
for f in 0 2048 4; do echo "== extra flags $f"; CTN_T32_EXTRA=$f timeout 300 python tools/t32_trace.py 10 2>&1 | tail -11; done
echo "== case 11/21 (MN-major B)"; timeout 300 python tools/t32_trace.py 11 2>&1 | tail -11
for f in 0 2048; do echo "== CTN_T32_DEBUG=$f"; CTN_T32_DEBUG=$f timeout 300 python tools/t32_check.py 10 11 12 2>&1 | grep "TFLOP\|OUT OF"; done
