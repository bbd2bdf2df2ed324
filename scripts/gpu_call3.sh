#!/bin/bash
# A/B of the round-2 kernel changes: tcgen05 attention v3 (decoupled Q ring), DWConv GELU on the MUFU pipe, tensor-core MS-TCN layers
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_mstcn_gpu.py tests/test_trans_head_gpu.py -m gpu -x -q > $O/pytest_c_kernels.log 2>&1; echo "kernels rc=$?"; tail -3 $O/pytest_c_kernels.log
timeout 900 python -m pytest tests/test_evp_gpu.py tests/test_job_gpu.py -m gpu -x -q -s > $O/pytest_c_evp.log 2>&1; echo "evp rc=$?"; tail -3 $O/pytest_c_evp.log; grep "chain\] argmax\|ref_init feats vs\|stress feats vs" $O/pytest_c_evp.log
for g in 0 1 2; do SURGVID_DW_GELU=$g REPS=20 python scripts/op_bench.py dwconv 2>&1 | sed "s/^/gelu$g /"; done | tee $O/dwconv_gelu_ab.log
for a in 0 1; do SURGVID_ATTN_TC=$a REPS=20 python scripts/op_bench.py attn 2>&1 | sed "s/^/tc$a /"; done | tee $O/attn_tc_ab.log
for m in 0 1; do SURGVID_MSTCN_TC=$m REPS=20 python scripts/mstcn_bench.py 2>&1 | tail -1 | sed "s/^/tc$m /"; done | tee $O/mstcn_tc_ab.log
SURGVID_PROFILE_CSV=$O/profile_ops_c.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c_default.json 2> $O/bench_c_default.err; echo "bench rc=$?"
SURGVID_DW_GELU=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_c_gelu0.json 2>/dev/null
SURGVID_DW_GELU=2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_c_gelu2.json 2>/dev/null
timeout 600 python bench.py --hw 480x854 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c_480.json 2> $O/bench_c_480.err; echo "480 rc=$?"
python - <<'PY'
import json
for f in ['bench_c_default','bench_c_gelu0','bench_c_gelu2','bench_c_480']:
    try:
        d=json.loads(open(f'gpurun_out/r02/{f}.json').read().strip().splitlines()[-1])
        k=d['kernel_classes']
        print(f, round(d['value']), round(d['ms_per_step'],2), d['e2e'] and round(d['e2e']['value']), 'attn', round(k['attention']['ms'],2), 'dw', round(k['dwconv3x3_gelu']['ms'],2), 'gemm', round(k['gemm_tcgen05']['ms'],2), d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
