#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
for c in 4 2; do SURGVID_DW_CPL=$c REPS=20 python scripts/op_bench.py dwconv 2>&1 | sed "s/^/cpl$c /"; done | tee $O/dwconv_cpl_ab.log
REPS=20 python scripts/op_bench.py stem 2>&1 | tee $O/stem_bench.log
for c in 4 2; do SURGVID_DW_CPL=$c timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "dwconv or stem" 2>&1 | tail -1 | sed "s/^/cpl$c /"; done
SURGVID_DW_CPL=2 timeout 900 python -m pytest tests/test_evp_gpu.py -m gpu -x -q -s -k "golden or ragged" > $O/pytest_f_evp.log 2>&1; echo "evp(cpl2) rc=$?"; grep "parity\] ref_init feats vs\|parity\] stress feats vs" $O/pytest_f_evp.log; tail -1 $O/pytest_f_evp.log
for c in 4 2; do SURGVID_DW_CPL=$c python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_f_cpl$c.json 2>/dev/null; done
python - <<'PY'
import json
for f in ['bench_f_cpl4','bench_f_cpl2']:
    try:
        d=json.loads(open(f'gpurun_out/r02/{f}.json').read().strip().splitlines()[-1])
        k=d['kernel_classes']
        print(f, round(d['value']), round(d['ms_per_step'],2), 'attn', round(k['attention']['ms'],2), 'dw', round(k['dwconv3x3_gelu']['ms'],2), 'gemm', round(k['gemm_tcgen05']['ms'],2), 'stem', round(k['stem_conv']['ms'],2), d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
