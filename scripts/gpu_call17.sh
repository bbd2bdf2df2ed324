#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
{
for e in 0 3 7 11 4 8; do echo "== EXP=$e (1 no B loads, 2 no A loads, 4 no result stores, 8 no epilogue work)"; SURGVID_GEMM_EXP=$e REPS=10 python scripts/gemm_bench.py 10,13,7,0 2>&1 | grep -v mbarrier; done
} > $O/gemm_exp2.log 2>&1
cat $O/gemm_exp2.log
