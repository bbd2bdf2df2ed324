#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python scripts/op_bench.py all > gpurun_out/op_bench.log 2>&1; cat gpurun_out/op_bench.log
REPS=1 python scripts/op_bench.py dwconv > /dev/null 2>&1 && \
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv -s 2 -c 1 -o gpurun_out/prof_dwconv -f python scripts/op_bench.py dwconv > gpurun_out/ncu_dw.log 2>&1
echo "ncu dw rc=$?"
REPS=1 python scripts/op_bench.py attn > /dev/null 2>&1 && \
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention -s 2 -c 1 -o gpurun_out/prof_attn -f python scripts/op_bench.py attn > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn rc=$?"
ls -la gpurun_out/*.ncu-rep
