#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k attention > $O/pytest_d_attn.log 2>&1; echo "attn rc=$?"; tail -2 $O/pytest_d_attn.log | cut -c1-200
if grep -q "mbarrier timeout" $O/pytest_d_attn.log; then echo "ATTN TC DEADLOCK -> falling back to SURGVID_ATTN_TC=0 for the rest"; export SURGVID_ATTN_TC=0; fi
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_mstcn_gpu.py tests/test_trans_head_gpu.py -m gpu -q -k "not attention" > $O/pytest_d_kernels.log 2>&1; echo "kernels rc=$?"; tail -3 $O/pytest_d_kernels.log | cut -c1-200
for g in 1 2; do SURGVID_DW_GELU=$g timeout 600 python -m pytest tests/test_evp_gpu.py -m gpu -q -s -k "golden or ragged" 2>&1 | grep "parity\]" | sed "s/^/gelu$g /"; done | tee $O/dwconv_gelu_parity.log
timeout 900 python -m pytest tests/test_evp_gpu.py tests/test_job_gpu.py -m gpu -x -q -s > $O/pytest_d_evp.log 2>&1; echo "evp rc=$?"; tail -3 $O/pytest_d_evp.log | cut -c1-200; grep "chain\] argmax" $O/pytest_d_evp.log
for g in 0 1 2; do SURGVID_DW_GELU=$g REPS=20 python scripts/op_bench.py dwconv 2>&1 | sed "s/^/gelu$g /"; done | tee $O/dwconv_gelu_ab.log
for a in 0 1; do SURGVID_ATTN_TC=$a REPS=20 python scripts/op_bench.py attn 2>&1 | grep -v mbarrier | sed "s/^/tc$a /"; done | tee $O/attn_tc_ab.log
SURGVID_PROFILE_CSV=$O/profile_ops_d.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_d_default.json 2> $O/bench_d_default.err; echo "bench rc=$?"
SURGVID_DW_GELU=2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_d_gelu2.json 2>/dev/null
SURGVID_ATTN_TC=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_d_attn0.json 2>/dev/null
python - <<'PY'
import json
for f in ['bench_d_default','bench_d_gelu2','bench_d_attn0']:
    try:
        d=json.loads(open(f'gpurun_out/r02/{f}.json').read().strip().splitlines()[-1])
        k=d['kernel_classes']
        print(f, round(d['value']), round(d['ms_per_step'],2), d['e2e'] and round(d['e2e']['value']), 'attn', round(k['attention']['ms'],2), 'dw', round(k['dwconv3x3_gelu']['ms'],2), 'gemm', round(k['gemm_tcgen05']['ms'],2), d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
