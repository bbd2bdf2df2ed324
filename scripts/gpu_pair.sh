#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== pair first contact"
SURGVID_GEMM_PAIR=1 timeout 120 python scripts/first_contact.py > gpurun_out/pair_first.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pair_first.log
echo "=== pair tests"
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "pair" > gpurun_out/pair_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pair_tests.log
echo "=== gemm bench pair=0 / pair=1"
SURGVID_GEMM_PAIR=0 timeout 120 python scripts/gemm_bench.py 3,4,5,6,8 2>&1 | tail -6
SURGVID_GEMM_PAIR=1 timeout 120 python scripts/gemm_bench.py 3,4,5,6,8 2>&1 | tail -6
