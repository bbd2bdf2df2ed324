#!/bin/bash
# same-box A/B of two builds of the native library: previous (commit f734fa4, before the elect.sync / pair-mode work) vs current
O=gpurun_out/r02; mkdir -p $O
L=$PWD/deep-learning-for-surgical-video-analysis_b200/lib
for rep in 1 2; do
for b in prev cur; do
  if [ $b = prev ]; then export SURGVID_LIB=$L/libsurgvid_prev.so; else unset SURGVID_LIB; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/ab_build_${b}_$rep.json 2>/dev/null; echo "$b $rep rc=$?"
done; done
unset SURGVID_LIB
python - <<'PY'
import json
for rep in (1,2):
  for b in ("prev","cur"):
    d=json.loads(open(f"gpurun_out/r02/ab_build_{b}_{rep}.json").read().strip().splitlines()[-1]); print(b, rep, round(d["value"]), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], {k:round(v["ms"],2) for k,v in d["kernel_classes"].items() if v["ms"]>1}, round(d['roofline']['frac'],3), round(d['roofline']['tensor']['frac'],3))
PY
