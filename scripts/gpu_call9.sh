#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests/test_trans_head_gpu.py tests/test_mstcn_gpu.py -m gpu -x -q -s > $O/pytest_h_trans.log 2>&1; echo "trans rc=$?"; grep "parity\]" $O/pytest_h_trans.log; tail -3 $O/pytest_h_trans.log | cut -c1-300
REPS=10 python scripts/mstcn_bench.py 2>&1 | tail -2 | tee $O/mstcn_head_bench.log
