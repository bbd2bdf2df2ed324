#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm" > $O/pytest_gemm_pair2.log 2>&1; echo "gemm tests rc=$?"; tail -2 $O/pytest_gemm_pair2.log | cut -c1-300
{
for pr in 0 1; do echo "== PAIR=$pr"; SURGVID_GEMM_PAIR=$pr REPS=10 python scripts/gemm_bench.py 10,11,12,13,7,0,2,3,4,16,17,18 2>&1 | grep -v mbarrier; done
} > $O/gemm_pair_relaxed_arrive.log 2>&1
cat $O/gemm_pair_relaxed_arrive.log
SURGVID_GEMM_PAIR=1 python scripts/gemm_trace.py 156800 1280 320 0 0 5 2>&1 | cut -c1-200 | tee $O/gemm_trace_pair2.log
