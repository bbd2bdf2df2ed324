#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm" > $O/pytest_gemm_f2.log 2>&1; echo "gemm tests rc=$?"; tail -2 $O/pytest_gemm_f2.log | cut -c1-300
REPS=10 python scripts/gemm_bench.py 10,11,12,13,7,0,2 2>&1 | grep -v mbarrier | tee $O/gemm_f2.log
SURGVID_PROFILE_CSV=$O/profile_ops_f2.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/f2_bench.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/f2_bench.json").read().strip().splitlines()[-1]); print("bench", round(d["value"]), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], 'e2e', round(d['e2e']['value']), {k:round(v["ms"],2) for k,v in d["kernel_classes"].items()}, d['roofline']['frac'], d['roofline']['tensor']['frac'])
PY
