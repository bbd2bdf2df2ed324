#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu -x > $O/pytest_gpu_final2.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu_final2.log | cut -c1-200
timeout 600 python __graft_entry__.py smoke > $O/smoke_final2.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_final2.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_final2.json 2> $O/bench_final2.err; echo "bench rc=$?"; cut -c1-300 $O/bench_final2.json
