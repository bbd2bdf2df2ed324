#!/bin/bash
# attention_tc + dwconv: control warps converged + elect.sync
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "dwconv or attention" > $O/elect2_tests.txt 2>&1; echo "tests rc=$?"; tail -2 $O/elect2_tests.txt | cut -c1-200
{ REPS=20 python scripts/op_bench.py dwconv; REPS=20 python scripts/op_bench.py attn; B=1159 REPS=10 python scripts/op_bench.py dwconv;  B=1159 REPS=10 python scripts/op_bench.py attn; } > $O/elect2_ops.txt 2>&1; cat $O/elect2_ops.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/elect2_bench.json 2>/dev/null
python bench.py --hw 480x854 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/elect2_bench480.json 2>$O/elect2_bench480.err || tail -3 $O/elect2_bench480.err
python - <<'PY'
import json
for f in ("elect2_bench","elect2_bench480"):
    try:
        d=json.loads(open(f"gpurun_out/r02/{f}.json").read().strip().splitlines()[-1]); print(f, round(d["value"]), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], {k:round(v["ms"],2) for k,v in d["kernel_classes"].items()})
    except Exception as e: print(f, "ERR", e)
PY
