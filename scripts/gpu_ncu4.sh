#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python scripts/mstcn_bench.py 2>&1 | tail -1 | tee gpurun_out/mstcn_bench.log
REPS=1 python scripts/mstcn_bench.py > /dev/null 2>&1 && REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mstcn -s 60 -c 20 --csv --log-file gpurun_out/mstcn_launches.csv python scripts/mstcn_bench.py > /dev/null 2>&1
echo "ncu mstcn rc=$?"
REPS=1 python scripts/op_bench.py dwconv > /dev/null 2>&1 && \
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv -s 2 -c 1 -o gpurun_out/prof_dwconv_tma -f python scripts/op_bench.py dwconv > gpurun_out/ncu_dw.log 2>&1
echo "ncu dw rc=$?"
