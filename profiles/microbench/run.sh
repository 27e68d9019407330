#!/bin/bash
# builds here (no GPU needed) or on the box; run under gpurun:  bash profiles/microbench/run.sh
set -e
cd "$(dirname "$0")"
[ -x fpbench ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o fpbench fpbench.cu
./fpbench
