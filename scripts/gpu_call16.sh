#!/bin/bash
# GEMM with the converged-warp / elect.sync issue pattern: tests, per-shape timing (single and pair mode), whole step
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm" > $O/pytest_gemm_elect.log 2>&1; echo "gemm tests rc=$?"; tail -2 $O/pytest_gemm_elect.log | cut -c1-300
{
for pr in 0 1; do echo "== PAIR=$pr"; SURGVID_GEMM_PAIR=$pr REPS=10 python scripts/gemm_bench.py 10,11,12,13,7,0,2,4,16,17,18 2>&1 | grep -v mbarrier; done
} > $O/gemm_elect.log 2>&1
cat $O/gemm_elect.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/elect_bench.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/elect_bench.json").read().strip().splitlines()[-1]); print("bench", round(d["value"]), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], {k:round(v["ms"],2) for k,v in d["kernel_classes"].items()})
PY
