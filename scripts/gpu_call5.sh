#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k attention > $O/pytest_e_attn.log 2>&1; echo "attn rc=$?"; grep -v mbarrier $O/pytest_e_attn.log | tail -2 | cut -c1-200
if grep -q "mbarrier timeout" $O/pytest_e_attn.log; then echo "ATTN TC DEADLOCK"; grep mbarrier $O/pytest_e_attn.log | sort | uniq -c | head; exit 1; fi
timeout 900 python -m pytest tests/test_evp_gpu.py -m gpu -x -q -s -k "golden or ragged or full_size" > $O/pytest_e_evp.log 2>&1; echo "evp rc=$?"; tail -2 $O/pytest_e_evp.log | cut -c1-200
for a in 0 1; do SURGVID_ATTN_TC=$a REPS=20 python scripts/op_bench.py attn 2>&1 | grep -v mbarrier | sed "s/^/tc$a /"; done | tee $O/attn_tc_ab.log
SURGVID_PROFILE_CSV=$O/profile_ops_e.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_e_default.json 2> $O/bench_e_default.err; echo "bench rc=$?"
timeout 600 python bench.py --hw 480x854 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_e_480.json 2> $O/bench_e_480.err; echo "480 rc=$?"
python - <<'PY'
import json
for f in ['bench_e_default','bench_e_480']:
    try:
        d=json.loads(open(f'gpurun_out/r02/{f}.json').read().strip().splitlines()[-1])
        k=d['kernel_classes']
        print(f, round(d['value']), round(d['ms_per_step'],2), d['e2e'] and round(d['e2e']['value']), 'attn', round(k['attention']['ms'],2), 'dw', round(k['dwconv3x3_gelu']['ms'],2), 'gemm', round(k['gemm_tcgen05']['ms'],2), 'ln', round(k['layernorm']['ms'],2), d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
