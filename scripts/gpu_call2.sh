#!/bin/bash
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k attention > gpurun_out/r02/pytest_attn_tc.log 2>&1; echo "attn rc=$?"; tail -15 gpurun_out/r02/pytest_attn_tc.log
timeout 900 python -m pytest tests/test_evp_gpu.py tests/test_job_gpu.py -m gpu -x -q -s > gpurun_out/r02/pytest_evp_b.log 2>&1; echo "evp rc=$?"; tail -5 gpurun_out/r02/pytest_evp_b.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02/bench_b_default.json 2> gpurun_out/r02/bench_b_default.err; echo "bench rc=$?"
SURGVID_ATTN_TC=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02/bench_b_noattntc.json 2>/dev/null; echo "bench rc=$?"
timeout 600 python bench.py --hw 480x854 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02/bench_b_480.json 2> gpurun_out/r02/bench_b_480.err; echo "480 rc=$?"
python - <<'PY'
import json
for f in ['bench_b_default','bench_b_noattntc','bench_b_480']:
    try:
        d=json.loads(open(f'gpurun_out/r02/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],2), d['e2e'] and round(d['e2e']['value']), 'attn', round(d['kernel_classes']['attention']['ms'],2), d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
