#!/bin/bash
# Round-end evidence run: full GPU test suite, smoke, default bench (+reference arm), ncu launch list and GEMM DRAM traffic.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; echo "=== $name" ; timeout "${TMO:-900}" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -n "${TAILN:-4}" gpurun_out/$name.log; return $rc; }
run pytest_gpu python -m pytest tests -q -m gpu -x || exit 1
run smoke python __graft_entry__.py smoke || exit 1
echo "=== bench (defaults)"
SURGVID_PROFILE_CSV=gpurun_out/profile_ops_final.csv timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?"; tail -c 300 gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py --batch 200 --micro-batch 200 --no-cpu-baseline > gpurun_out/bench_batch200.json 2> gpurun_out/bench_batch200.err; echo "b200 rc=$?"
timeout 600 python bench.py --fold-head 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_foldhead.json 2> gpurun_out/bench_foldhead.err; echo "fold rc=$?"
python scripts/mstcn_bench.py 2>&1 | tail -1 | tee gpurun_out/mstcn_bench.log
if [ "${NCU:-1}" = "1" ]; then
CMD="python bench.py --frames 800 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1540 -c 365 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_final.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_final.csv
fi
