#!/bin/bash
# A/B of the DRAM-locality work orders (DWConv channel block fastest, attention head fastest), then the affected tests and a short bench.
mkdir -p gpurun_out/r02
{
for o in 0 1; do
  echo "== DW_ORDER=$o ATTN_ORDER=$o B=200"; SURGVID_ATTN_ORDER=$o REPS=20 python scripts/op_bench.py dwconv
  SURGVID_ATTN_ORDER=$o REPS=20 python scripts/op_bench.py attn
  echo "== DW_ORDER=$o ATTN_ORDER=$o B=1159"; SURGVID_ATTN_ORDER=$o B=1159 REPS=10 python scripts/op_bench.py dwconv
  SURGVID_ATTN_ORDER=$o B=1159 REPS=10 python scripts/op_bench.py attn
done
} > gpurun_out/r02/order_ab.txt 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "dwconv or attention" > gpurun_out/r02/order_tests.txt 2>&1; echo "tests rc=$?"
for o in "0 0" "1 0" "0 1" "1 1"; do set -- $o
  SURGVID_DW_ORDER=$1 SURGVID_ATTN_ORDER=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02/order_bench_$1$2.json 2> gpurun_out/r02/order_bench_$1$2.err; echo "bench $1 $2 rc=$?"
done
tail -3 gpurun_out/r02/order_tests.txt
cat gpurun_out/r02/order_ab.txt
python - <<'PY'
import json
for k in ("00","10","01","11"):
    try:
        d=json.loads(open(f"gpurun_out/r02/order_bench_{k}.json").read().strip().splitlines()[-1]); print(k, d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])
    except Exception as e: print(k, "ERR", e)
PY
