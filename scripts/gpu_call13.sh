#!/bin/bash
# GEMM: which operand stream bounds the stage-3 GEMMs (timing-only experiments, wrong results) + residual L2 prefetch A/B
O=gpurun_out/r02; mkdir -p $O
{
for e in 0 1 2 3; do echo "== EXP=$e (1: no B loads, 2: no A loads)"; SURGVID_GEMM_EXP=$e REPS=10 python scripts/gemm_bench.py 10,11,12,13,7,0 2>&1 | grep -v mbarrier; done
for r in 0 1; do echo "== RES_PREFETCH=$r"; SURGVID_GEMM_RES_PREFETCH=$r REPS=10 python scripts/gemm_bench.py 11,12,2,4,17 2>&1 | grep -v mbarrier; done
} > $O/gemm_exp.log 2>&1
cat $O/gemm_exp.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm" > $O/pytest_gemm_rpf.log 2>&1; echo "gemm tests rc=$?"; tail -2 $O/pytest_gemm_rpf.log | cut -c1-200
for r in 0 1; do SURGVID_GEMM_RES_PREFETCH=$r python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/rpf_bench_$r.json 2>/dev/null; done
python - <<'PY'
import json
for k in ("0","1"):
    d=json.loads(open(f"gpurun_out/r02/rpf_bench_{k}.json").read().strip().splitlines()[-1]); print("res_prefetch",k, round(d["value"]), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], round(d["kernel_classes"]["gemm_tcgen05"]["ms"],2))
PY
