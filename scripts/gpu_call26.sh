#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "dwconv" > $O/dw_revert_tests.txt 2>&1; echo "tests rc=$?"; tail -1 $O/dw_revert_tests.txt
run() { name=$1; shift; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/ab_$name.json 2>/dev/null; echo "$name rc=$?"; }
run cur_a X=1
run bufs2 SURGVID_GEMM_EPI_BUFS=2
run bufs3 SURGVID_GEMM_EPI_BUFS=3
run lag2 SURGVID_GEMM_EPI_LAG=2
run cur_b X=1
python - <<'PY'
import json
for b in ("cur_a","bufs2","bufs3","lag2","cur_b"):
    d=json.loads(open(f"gpurun_out/r02/ab_{b}.json").read().strip().splitlines()[-1]); print(b, round(d["value"]), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], {k:round(v["ms"],2) for k,v in d["kernel_classes"].items() if v["ms"]>1}, round(d['roofline']['frac'],3))
PY
