#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm" > $O/pytest_i_gemm.log 2>&1; echo "gemm tests rc=$?"; grep -v mbarrier $O/pytest_i_gemm.log | tail -2 | cut -c1-200
for pr in 0 1; do SURGVID_GEMM_PAIR=$pr REPS=10 python scripts/gemm_bench.py 10,11,12,13,3,4,17,18 2>&1 | grep -v mbarrier | sed "s/^/pair$pr /"; done | tee $O/gemm_pair_ab_r02.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_i_default.json 2>/dev/null; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_i_default.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']),'fp32',round(d['e2e']['from_fp32_tensors']['value']),d['clocks']['sm_mhz'])
PY
