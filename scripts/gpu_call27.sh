#!/bin/bash
# final state of the round: full GPU suite, smoke, the driver-style bench, then previous build vs final build on this same box
O=gpurun_out/r02; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu -x > $O/pytest_gpu_final6.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu_final6.log | cut -c1-200
timeout 600 python __graft_entry__.py smoke > $O/smoke_final6.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_final6.log
SURGVID_PROFILE_CSV=$O/profile_ops_final6.csv timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_final6.json 2> $O/bench_final6.err; echo "bench rc=$?"
L=$PWD/deep-learning-for-surgical-video-analysis_b200/lib
SURGVID_LIB=$L/libsurgvid_prev.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/ab2_build_prev.json 2>/dev/null; echo "prev rc=$?"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/ab2_build_cur.json 2>/dev/null; echo "cur rc=$?"
python - <<'PY'
import json
for f in ("bench_final6","ab2_build_prev","ab2_build_cur"):
    d=json.loads(open(f"gpurun_out/r02/{f}.json").read().strip().splitlines()[-1]); print(f, round(d["value"]), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], 'e2e', d.get('e2e') and round(d['e2e']['value']), {k:round(v["ms"],2) for k,v in d["kernel_classes"].items() if v["ms"]>1}, round(d['roofline']['frac'],3), round(d['roofline']['tensor']['frac'],3))
PY
