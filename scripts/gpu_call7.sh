#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
for cfg in "4 4 2" "4 4 3" "2 4 2" "2 6 4" "2 8 6" "2 8 4" "2 5 3"; do set -- $cfg; SURGVID_DW_CPL=$1 SURGVID_DW_STAGES=$2 SURGVID_DW_PREFETCH=$3 REPS=20 python scripts/op_bench.py dwconv 2>&1 | sed "s/^/cpl$1 st$2 pf$3 /"; done | tee $O/dwconv_ring_sweep.log
