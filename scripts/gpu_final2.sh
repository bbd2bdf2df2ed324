#!/bin/bash
# Round-end evidence run on ONE B200: full GPU test suite, smoke, default bench as the driver launches it (+ reference arm), batch
# variants, the 80-video job on one GPU, the 480x854 line, MS-TCN, and ncu (launch list of one micro-batch + --set full of the GEMM).
O=gpurun_out/r02; mkdir -p $O
cd "$(dirname "$0")/.."
run() { name=$1; shift; echo "=== $name" ; timeout "${TMO:-1200}" "$@" > $O/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -n "${TAILN:-3}" $O/$name.log | cut -c1-220; return $rc; }
run pytest_gpu_final5 python -m pytest tests -q -m gpu -x -s || exit 1
run smoke_final5 python __graft_entry__.py smoke || exit 1
echo "=== bench (as the driver runs it)"
SURGVID_PROFILE_CSV=$O/profile_ops_final5.csv timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_final5.json 2> $O/bench_final5.err; echo "rc=$?"; tail -c 300 $O/bench_final5.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/bench_final5_reference.json 2> /dev/null; echo "ref rc=$?"
timeout 600 python bench.py --batch 200 --micro-batch 200 --no-cpu-baseline --no-e2e > $O/bench_final5_batch200.json 2> /dev/null; echo "b200 rc=$?"
timeout 900 python bench.py --hw 480x854 --batch 64 --steps 10 --warmup 3 > $O/bench_final5_480.json 2> /dev/null; echo "480 rc=$?"
timeout 900 python bench.py --workload cholec80x80 --steps 1 --warmup 1 --no-cpu-baseline > $O/bench_final5_job_n1.json 2> /dev/null; echo "job n1 rc=$?"
REPS=20 python scripts/mstcn_bench.py 2>&1 | tail -1 | tee $O/mstcn_bench_final5.log
python - <<'PY'
import json
for f in ['bench_final5','bench_final5_batch200','bench_final5_batch1150','bench_final5_foldhead','bench_final5_480','bench_final5_job_n1','bench_final5_reference']:
    try:
        d=json.loads(open(f'gpurun_out/r02/{f}.json').read().strip().splitlines()[-1])
        k=d.get('kernel_classes') or {}
        print(f, round(d['value'],1), round(d['ms_per_step'],2), d.get('e2e') and round(d['e2e']['value']), {n:round(v['ms'],2) for n,v in k.items() if v['ms']>1}, d.get('roofline') and round(d['roofline']['frac'],3), d.get('clocks') and d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
if [ "${NCU:-1}" = "1" ]; then
NCUB="ncu --clock-control none"
export SURGVID_NCU_RANGE=1
CMD="python bench.py --frames 800 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > $O/ncu_plain_bench.log 2>&1 && $NCUB --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file $O/launches_r02_final5_batch800.csv $CMD > $O/ncu_launches.log 2>&1
echo "launch list rc=$?"
unset SURGVID_NCU_RANGE
for sh in 10 11; do
  CMD="python scripts/gemm_bench.py $sh"
  REPS=1 $CMD > $O/ncu_plain_gemm_$sh.log 2>&1 && REPS=1 $NCUB --set full --import-source on -k regex:gemm_bf16_tcgen05 -s 2 -c 1 -o $O/ncu_gemm_final5_shape$sh -f $CMD > $O/ncu_gemm_$sh.log 2>&1
  echo "gemm shape $sh rc=$?"
  ncu -i $O/ncu_gemm_final5_shape$sh.ncu-rep --page raw --csv > $O/ncu_gemm_final5_shape${sh}_raw.csv 2>/dev/null
done
rm -f $O/ncu_gemm_final5_shape11.ncu-rep
du -sh $O
fi
