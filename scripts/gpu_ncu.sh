#!/bin/bash
# round-2 ncu evidence (one GPU; every command runs plain first and must exit 0 before it runs under ncu)
O=gpurun_out/r02; mkdir -p $O
NCU="ncu --clock-control none"
# (1) launch list of one whole run of a single 800-frame video (3 warm-up steps + 1 timed + the per-class profile pass): per-launch device time and DRAM bytes
CMD="python bench.py --frames 800 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > $O/ncu_plain_bench.log 2>&1 && $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --csv --log-file $O/launches_r02_batch800.csv $CMD > $O/ncu_launches.log 2>&1
echo "launch list rc=$?"
# (2) --set full of the GEMM at the stage-3 fc1 / fc2cat / residual proj shapes and one stage-1 K=64 shape (gemm_bench indices 10,11,12,0)
CMD="python scripts/gemm_bench.py 10,11,12,0"
REPS=1 $CMD > $O/ncu_plain_gemm.log 2>&1 && REPS=1 $NCU --set full --import-source on -k regex:gemm_bf16_tcgen05 -c 12 -o $O/ncu_gemm_r02 -f $CMD > $O/ncu_gemm.log 2>&1
echo "gemm rc=$?"
# (3) tcgen05 attention (stage 3 / 1 / 2 / 4 shapes at 200 frames) and the mma.sync cross-attention
CMD="python scripts/op_bench.py attn"
REPS=1 $CMD > $O/ncu_plain_attn.log 2>&1 && REPS=1 $NCU --set full --import-source on -k regex:attention -c 15 -o $O/ncu_attn_r02 -f $CMD > $O/ncu_attn.log 2>&1
echo "attn rc=$?"
# (4) DWConv3x3+GELU (stage 3 / 1 / 2 / 4 shapes at 200 frames)
CMD="python scripts/op_bench.py dwconv"
REPS=1 $CMD > $O/ncu_plain_dw.log 2>&1 && REPS=1 $NCU --set full --import-source on -k regex:dwconv3x3 -c 12 -o $O/ncu_dwconv_r02 -f $CMD > $O/ncu_dw.log 2>&1
echo "dwconv rc=$?"
# (5) MS-TCN over the 80 sequences: stage-1 projection + tensor-core layer kernel
CMD="python scripts/mstcn_bench.py"
REPS=1 $CMD > $O/ncu_plain_mstcn.log 2>&1 && REPS=1 $NCU --set full --import-source on -k regex:mstcn -s 60 -c 20 -o $O/ncu_mstcn_r02 -f $CMD > $O/ncu_mstcn.log 2>&1
echo "mstcn rc=$?"
ls -la $O/*.ncu-rep
