#!/bin/bash
# iteration loop: kernel parity, model parity, then a profiled bench at one micro-batch
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; echo "=== $name" ; timeout "${TMO:-900}" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -n "${TAILN:-6}" gpurun_out/$name.log; return $rc; }
run kern python -m pytest tests/test_kernels_gpu.py -q -x || exit 1
TAILN=25 run evp python -m pytest tests/test_evp_gpu.py -q -s -x || exit 1
run mstcn python -m pytest tests/test_mstcn_gpu.py -q -x || exit 1
for mb in ${MBS:-200}; do
  SURGVID_PROFILE_CSV=gpurun_out/profile_ops_mb$mb.csv timeout 600 python bench.py --steps 2 --warmup 3 --micro-batch $mb --batch $( [ "${BATCH_EQ_MB:-0}" = "1" ] && echo $mb || echo ${BATCH:-200} ) --no-e2e --no-cpu-baseline > gpurun_out/bench_mb$mb.json 2> gpurun_out/bench_mb$mb.err
  echo "bench mb=$mb rc=$?"; tail -c 300 gpurun_out/bench_mb$mb.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_mb$mb.json").read().strip().splitlines()[-1])
    print("mb", $mb, "frames/s %.1f"%d["value"], "ms/step %.1f"%d["ms_per_step"], "gemm TF/s %.1f"%d["roofline"]["achieved"], "launches", d["gpu_launches"])
    for k,v in d["kernel_classes"].items(): print("   ", k, "%.2f ms"%v["ms_per_step"], "share %.3f"%v["share"], v["launches_per_step"])
except Exception as e: print("parse fail", e)
PY
done
