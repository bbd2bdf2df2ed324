#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "stem or gauss" > $O/pytest_j.log 2>&1; echo "stem/gauss tests rc=$?"; tail -2 $O/pytest_j.log | cut -c1-200
timeout 900 python -m pytest tests/test_evp_gpu.py -m gpu -x -q -s -k "golden or ragged or 480 or variants" > $O/pytest_j_evp.log 2>&1; echo "evp rc=$?"; tail -1 $O/pytest_j_evp.log; grep "parity\] ref_init feats vs\|parity\] stress feats vs" $O/pytest_j_evp.log
REPS=20 python scripts/op_bench.py stem 2>&1 | tee $O/stem_bench2.log
REPS=20 python scripts/op_bench.py im2col 2>&1 | grep gauss | tee -a $O/stem_bench2.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_j.json 2>/dev/null; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_j.json').read().strip().splitlines()[-1])
k=d['kernel_classes']
print('value',round(d['value']),'ms',round(d['ms_per_step'],2),{n:round(v['ms'],2) for n,v in k.items() if v['ms']>1},d['clocks']['sm_mhz'])
PY
