#!/bin/bash
# GEMM TMA epilogues: staging buffers per warp (2 / 4) and, for the residual form, the store lag (1..3)
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm" > $O/pytest_gemm_epi.log 2>&1; echo "gemm tests rc=$?"; tail -2 $O/pytest_gemm_epi.log | cut -c1-300
{
for cfg in "2 1" "4 1" "4 2" "4 3" "3 1"; do set -- $cfg
  echo "== EPI_BUFS=$1 EPI_LAG=$2"; SURGVID_GEMM_EPI_BUFS=$1 SURGVID_GEMM_EPI_LAG=$2 REPS=10 python scripts/gemm_bench.py 10,11,12,13,7,0,2,4,17,18 2>&1 | grep -v mbarrier
done
} > $O/gemm_epi_bufs_ab.log 2>&1
cat $O/gemm_epi_bufs_ab.log
for cfg in "2 1" "4 1" "4 2"; do set -- $cfg
  SURGVID_GEMM_EPI_BUFS=$1 SURGVID_GEMM_EPI_LAG=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/epi_bench_$1$2.json 2>/dev/null
done
python - <<'PY'
import json
for k in ("21","41","42"):
    d=json.loads(open(f"gpurun_out/r02/epi_bench_{k}.json").read().strip().splitlines()[-1]); print("bufs/lag",k, round(d["value"]), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], round(d["kernel_classes"]["gemm_tcgen05"]["ms"],2))
PY
