#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
REPS=1 python scripts/gemm_bench.py 1,3 > gpurun_out/gemm_plain.log 2>&1 && \
REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 4 -c 4 -o gpurun_out/prof_gemm2 -f python scripts/gemm_bench.py 1,3 > gpurun_out/ncu_gemm2.log 2>&1
echo "ncu gemm rc=$?"
CMD="python bench.py --frames 200 --steps 1 --warmup 3 --micro-batch 200 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dwconv -s 150 -c 2 -o gpurun_out/prof_dwconv -f $CMD > gpurun_out/ncu_dw.log 2>&1
echo "ncu dwconv rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 160 -c 2 -o gpurun_out/prof_attn -f $CMD > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn rc=$?"
ls -la gpurun_out/*.ncu-rep
