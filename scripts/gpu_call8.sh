#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > $O/pytest_g_kernels.log 2>&1; echo "kernels rc=$?"; grep -v mbarrier $O/pytest_g_kernels.log | tail -3 | cut -c1-200
if grep -q "mbarrier timeout" $O/pytest_g_kernels.log; then echo DEADLOCK; grep mbarrier $O/pytest_g_kernels.log | sort | uniq -c | head -5; export SURGVID_GEMM_TMA_OUT=0; fi
timeout 900 python -m pytest tests/test_evp_gpu.py -m gpu -x -q -s > $O/pytest_g_evp.log 2>&1; echo "evp rc=$?"; tail -2 $O/pytest_g_evp.log | cut -c1-200; grep "parity\] ref_init feats vs\|parity\] stress feats vs" $O/pytest_g_evp.log
for t in 1 0; do SURGVID_GEMM_TMA_OUT=$t REPS=10 python scripts/gemm_bench.py 2,4,6,11,12,8,17 2>&1 | sed "s/^/tma$t /"; done | tee $O/gemm_tma_out_ab.log
SURGVID_PROFILE_CSV=$O/profile_ops_g.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_g_default.json 2> $O/bench_g_default.err; echo "bench rc=$?"
SURGVID_GEMM_TMA_OUT=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_g_tma0.json 2>/dev/null
python - <<'PY'
import json
for f in ['bench_g_default','bench_g_tma0']:
    try:
        d=json.loads(open(f'gpurun_out/r02/{f}.json').read().strip().splitlines()[-1])
        k=d['kernel_classes']
        print(f, round(d['value']), round(d['ms_per_step'],2), d['e2e'] and round(d['e2e']['value']), 'attn', round(k['attention']['ms'],2), 'dw', round(k['dwconv3x3_gelu']['ms'],2), 'gemm', round(k['gemm_tcgen05']['ms'],2), 'roof', round(d['roofline']['frac'],3), round(d['roofline']['tensor']['frac'],3), d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
