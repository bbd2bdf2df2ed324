#!/bin/bash
# round-2 ncu evidence (one GPU; every command runs plain first and must exit 0 before it runs under ncu).
# gpurun copies back at most 64 MiB: a report with source costs ~2.2 MB per kernel instance, so every capture takes few instances.
O=gpurun_out/r02; mkdir -p $O
NCU="ncu --clock-control none"
# (1) launch list of ONE 800-frame micro-batch + MS-TCN over its 800 features (the per-class profile pass of bench.py, bracketed with
#     cudaProfilerStart/Stop): per-launch device time and DRAM bytes
export SURGVID_NCU_RANGE=1
CMD="python bench.py --frames 800 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > $O/ncu_plain_bench.log 2>&1 && $NCU --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file $O/launches_r02_batch800.csv $CMD > $O/ncu_launches.log 2>&1
echo "launch list rc=$?"
unset SURGVID_NCU_RANGE
# (2) --set full of the GEMM: stage-3 fc1 / fc2cat / residual proj and one stage-1 K=64 shape (gemm_bench indices 10,11,12,0): the third
#     launch of each shape (2 warm-ups before it)
i=0
for sh in 10 11 12 0; do
  CMD="python scripts/gemm_bench.py $sh"
  REPS=1 $CMD > $O/ncu_plain_gemm_$sh.log 2>&1 && REPS=1 $NCU --set full --import-source on -k regex:gemm_bf16_tcgen05 -s 2 -c 1 -o $O/ncu_gemm_r02_shape$sh -f $CMD > $O/ncu_gemm_$sh.log 2>&1
  echo "gemm shape $sh rc=$?"
done
# (3) tcgen05 attention: stage 3 / 1 / 2 / 4 shapes (200 frames) + the mma.sync cross-attention: third launch of each
CMD="python scripts/op_bench.py attn"
REPS=1 $CMD > $O/ncu_plain_attn.log 2>&1 && REPS=1 $NCU --set full --import-source on -k regex:attention -c 15 --launch-skip 0 -o $O/ncu_attn_tmp -f $CMD > $O/ncu_attn.log 2>&1
echo "attn rc=$?"
# keep only launches 3, 6 (stage 3 and stage 1: ids are 1-based in the report) by re-exporting the raw page; the report itself is dropped if large
ncu -i $O/ncu_attn_tmp.ncu-rep --page raw --csv > $O/ncu_attn_r02_raw.csv 2>/dev/null
ncu -i $O/ncu_attn_tmp.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:attention_tc_single --launch-skip 2 --launch-count 1 > $O/ncu_attn_r02_source_stage3.csv 2>/dev/null
rm -f $O/ncu_attn_tmp.ncu-rep
# (4) DWConv3x3+GELU stage-3 and stage-1 shapes (third launch of the first two shapes)
CMD="python scripts/op_bench.py dwconv"
REPS=1 $CMD > $O/ncu_plain_dw.log 2>&1 && REPS=1 $NCU --set full --import-source on -k regex:dwconv3x3 -c 6 -o $O/ncu_dw_tmp -f $CMD > $O/ncu_dw.log 2>&1
echo "dwconv rc=$?"
ncu -i $O/ncu_dw_tmp.ncu-rep --page raw --csv > $O/ncu_dwconv_r02_raw.csv 2>/dev/null
ncu -i $O/ncu_dw_tmp.ncu-rep --page source --csv --print-source cuda,sass --launch-skip 2 --launch-count 1 > $O/ncu_dwconv_r02_source_stage3.csv 2>/dev/null
rm -f $O/ncu_dw_tmp.ncu-rep
# (5) MS-TCN over the 80 sequences: raw metrics of every kernel of one forward (20 launches)
CMD="python scripts/mstcn_bench.py"
REPS=1 $CMD > $O/ncu_plain_mstcn.log 2>&1 && REPS=1 $NCU --set full -k regex:mstcn -s 60 -c 20 -o $O/ncu_mstcn_tmp -f $CMD > $O/ncu_mstcn.log 2>&1
echo "mstcn rc=$?"
ncu -i $O/ncu_mstcn_tmp.ncu-rep --page raw --csv > $O/ncu_mstcn_r02_raw.csv 2>/dev/null
rm -f $O/ncu_mstcn_tmp.ncu-rep
du -sh $O; ls -la $O | head -40
