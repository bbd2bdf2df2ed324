#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
{ SURGVID_GEMM_PAIR=1 python scripts/gemm_trace.py 156800 1280 320 0 0 6; SURGVID_GEMM_PAIR=0 python scripts/gemm_trace.py 156800 1280 320 0 0 6; } > $O/gemm_trace_epi.log 2>&1; cat $O/gemm_trace_epi.log | cut -c1-200,380-900
