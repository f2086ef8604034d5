#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_r2_m.sh $1
