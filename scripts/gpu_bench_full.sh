#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
MB=${MB:-200}
echo "=== full bench mb=$MB"
SURGVID_PROFILE_CSV=gpurun_out/profile_ops.csv timeout 900 python bench.py --steps 3 --warmup 3 --micro-batch $MB > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "rc=$?"; tail -c 600 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json | head -c 3000; echo
echo "=== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json | head -c 1500; echo
if [ "${NCU:-1}" = "1" ]; then
echo "=== ncu launch list"
CMD="python bench.py --frames 400 --steps 1 --warmup 3 --micro-batch $MB --no-e2e --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_run.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_run.log; wc -l gpurun_out/launches.csv
fi
