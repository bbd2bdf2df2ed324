#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export SURGVID_GEMM_PAIR=0
REPS=1 python scripts/gemm_bench.py 10,11 > gpurun_out/gemm_plain.log 2>&1 && \
REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 4 -c 2 -o gpurun_out/prof_gemm3 -f python scripts/gemm_bench.py 10,11 > gpurun_out/ncu_gemm3.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_gemm3.log
