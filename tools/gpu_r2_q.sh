#!/bin/bash
mkdir -p gpurun_out
CTN_T32_TRACE_LAUNCH=0 timeout 300 python tools/t32_trace.py 21 > gpurun_out/q_trace21_first.log 2>&1; tail -7 gpurun_out/q_trace21_first.log | cut -c1-220
timeout 600 python tools/t32_check.py 0 1 2 5 6 9 12 13 15 21 22 2>&1 | grep "OUT OF\|done,\|TFLOP\|flags [1-9]"
CTN_TF32_NO_TMA=1 timeout 600 python tools/t32_check.py 0 1 2 3 4 5 6 2>&1 | grep "OUT OF\|done,\|flags [1-9]"
