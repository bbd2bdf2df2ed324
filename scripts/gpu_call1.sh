#!/bin/bash
# round-2 call 1: full GPU test suite, default bench, the 80-video job on 1 GPU (reduced), native 480x854 bench
mkdir -p gpurun_out/r02
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02/pytest_gpu_a.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/r02/pytest_gpu_a.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02/bench_a_default.json 2> gpurun_out/r02/bench_a_default.err; echo "bench default rc=$?"
timeout 900 python bench.py --workload cholec80x80 --videos 12 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02/bench_a_job12.json 2> gpurun_out/r02/bench_a_job12.err; echo "job12 rc=$?"
timeout 600 python bench.py --hw 480x854 --batch 64 --steps 5 --warmup 3 > gpurun_out/r02/bench_a_480.json 2> gpurun_out/r02/bench_a_480.err; echo "480 rc=$?"
cut -c1-400 gpurun_out/r02/bench_a_default.json; tail -5 gpurun_out/r02/bench_a_default.err
cut -c1-300 gpurun_out/r02/bench_a_job12.json; tail -5 gpurun_out/r02/bench_a_job12.err
cut -c1-300 gpurun_out/r02/bench_a_480.json; tail -5 gpurun_out/r02/bench_a_480.err
