#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for mb in ${MBS:-8 16 32 64 100 200}; do
  echo "=== micro_batch $mb"
  timeout 600 python bench.py --steps 2 --warmup 3 --micro-batch $mb --no-e2e --no-cpu-baseline > gpurun_out/bench_mb$mb.json 2> gpurun_out/bench_mb$mb.err
  echo "rc=$?"; tail -c 400 gpurun_out/bench_mb$mb.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_mb$mb.json").read().strip().splitlines()[-1])
    print("mb", $mb, "frames/s %.1f"%d["value"], "ms/step %.1f"%d["ms_per_step"], "gemm TF/s %.1f"%d["roofline"]["achieved"], "launches", d["gpu_launches"])
    for k,v in d["kernel_classes"].items(): print("   ", k, "%.2f ms"%v["ms_per_step"], "share %.3f"%v["share"], v["launches_per_step"])
except Exception as e: print("parse fail", e)
PY
done
