#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_evp_gpu.py -x -q -m gpu -k "dwconv or golden or evp" > $O/dw_restore_tests.txt 2>&1; echo "tests rc=$?"; tail -1 $O/dw_restore_tests.txt
L=$PWD/deep-learning-for-surgical-video-analysis_b200/lib
SURGVID_LIB=$L/libsurgvid_prev.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/ab3_build_prev.json 2>/dev/null; echo "prev rc=$?"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/ab3_build_cur.json 2>/dev/null; echo "cur rc=$?"
python - <<'PY'
import json
for f in ("ab3_build_prev","ab3_build_cur"):
    d=json.loads(open(f"gpurun_out/r02/{f}.json").read().strip().splitlines()[-1]); print(f, round(d["value"]), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], {k:round(v["ms"],2) for k,v in d["kernel_classes"].items() if v["ms"]>1}, round(d['roofline']['frac'],3), round(d['roofline']['tensor']['frac'],3))
PY
