#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu -x > $O/pytest_gpu_final7.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu_final7.log | cut -c1-200
timeout 600 python __graft_entry__.py smoke > $O/smoke_final7.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_final7.log
SURGVID_PROFILE_CSV=$O/profile_ops_final7.csv timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_final7.json 2> $O/bench_final7.err; echo "bench rc=$?"; cut -c1-300 $O/bench_final7.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench_final7.json").read().strip().splitlines()[-1]); print("bench", round(d["value"]), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], 'e2e', round(d['e2e']['value']), {k:round(v["ms"],2) for k,v in d["kernel_classes"].items()}, d['roofline']['frac'], d['roofline']['tensor']['frac'])
PY
