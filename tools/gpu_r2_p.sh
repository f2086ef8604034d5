#!/bin/bash
# timing experiments on the tcgen05 kernel (CTN_T32_DEBUG flags; results are wrong by construction)
mkdir -p gpurun_out
for f in 0 4 8 256 512 204 460 716 1228 1740; do echo "== CTN_T32_DEBUG=$f"; CTN_T32_DEBUG=$f timeout 300 python tools/t32_check.py 10 12 2>&1 | grep TFLOP; done
