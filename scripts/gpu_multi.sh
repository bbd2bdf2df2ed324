#!/bin/bash
# usage: gpu_multi.sh N [steps] [warmup] — the 80-video job on N GPUs of one box (torchrun, one rank per GPU) + the multi-GPU tests
N=${1:-2}; K=${2:-2}; W=${3:-1}
O=gpurun_out/r02; mkdir -p $O
nvidia-smi -L | head -8
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_job_gpu.py -m gpu -x -q -s -k "sharded or second_gpu or cyclic" > $O/pytest_multi_n$N.log 2>&1; echo "multi-gpu tests rc=$?"; tail -3 $O/pytest_multi_n$N.log | cut -c1-200
fi
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps $K --warmup $W > $O/bench_job_n$N.json 2> $O/bench_job_n$N.err; echo "bench N=$N rc=$?"
tail -5 $O/bench_job_n$N.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/bench_job_ref_n$N.json 2> /dev/null; echo "ref rc=$?"
python - <<PY
import json
d=json.loads(open('$O/bench_job_n$N.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'scaling', d['scaling'], 'n_gpus', d['n_gpus'])
print('e2e', {k:(round(v) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k in ('value','h2d_bytes_per_step','d2h_bytes_per_step','steps','host_cores_bound')}, 'fp32', round(d['e2e']['from_fp32_tensors']['value']))
print({k:v for k,v in d['config'].items() if k in ('lpt_frames_per_rank','lpt_imbalance','gather','gather_verified')})
print('clocks', d['clocks'], 'launches', d['gpu_launches'])
r=json.loads(open('$O/bench_job_ref_n$N.json').read().strip().splitlines()[-1]); print('reference arm', round(r['value'],1), r['cpu_baseline']['cores'])
PY
