#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python scripts/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; echo "rc=$?"; cat gpurun_out/gemm_bench.log
REPS=1 python scripts/gemm_bench.py 0,3 > gpurun_out/gemm_plain.log 2>&1 && \
REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -c 6 -o gpurun_out/prof_gemm -f python scripts/gemm_bench.py 0,3 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/ncu_gemm.log; ls -la gpurun_out/*.ncu-rep
