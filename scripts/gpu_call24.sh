#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
{ for e in 0 1 2 4 5; do echo "== DW_EXP=$e (1 no math/stores, 2 no stores, 4 no TMA loads)"; SURGVID_DW_EXP=$e B=1159 REPS=10 python scripts/op_bench.py dwconv; done; } > $O/dwconv_skeleton.log 2>&1; cat $O/dwconv_skeleton.log
