#!/bin/bash
# Round-2 ncu evidence in one GPU-box visit: launch lists of the timed inference region (fp16x2, bf16x3 and bf16), of one training
# step and of one cancer fine-tune step; full captures of the kernels under work (edge forward ws, edge backward ws, node
# kernel, TMA GEMM); the tcgen05.mma timing probe.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
python scripts/umma_timing.py > gpurun_out/umma_timing.txt 2>&1; cat gpurun_out/umma_timing.txt
python scripts/prof_edge_bwd.py > gpurun_out/prof_edge_bwd.txt 2>&1; grep "edge_bwd" gpurun_out/prof_edge_bwd.txt
for prec in fp16x2 bf16x3 bf16; do
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches_${prec}.csv \
      python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline --profile --precision ${prec} > gpurun_out/ncu_bench_${prec}.log 2>&1
  echo "launch list ${prec} exit $?"
  ncu --set full --clock-control none --import-source on -k regex:edge_fwd_ws -s 2 -c 1 -f -o gpurun_out/prof_edge_ws_${prec} \
      python scripts/prof_edge.py ${prec} > gpurun_out/ncu_full_${prec}.log 2>&1
  echo "full capture ${prec} exit $?"
done
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_train.csv python scripts/prof_train.py bf16x3 train fused > gpurun_out/ncu_train.log 2>&1
echo "train launch list exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_cmp.csv python scripts/prof_train.py bf16x3 comparative fused > gpurun_out/ncu_cmp.log 2>&1
echo "comparative launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:edge_bwd_ws -s 8 -c 1 -f -o gpurun_out/prof_edge_bwd_ws \
    python scripts/prof_train.py bf16x3 train fused > gpurun_out/ncu_full_edge_bwd_ws.log 2>&1
echo "edge_bwd_ws full capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:node_post_pre_tc -s 27 -c 1 -f -o gpurun_out/prof_node \
    python scripts/prof_node.py > gpurun_out/ncu_full_node.log 2>&1
echo "node full capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tma_kernel -s 4 -c 2 -f -o gpurun_out/prof_gemm \
    python scripts/prof_train.py bf16x3 train fused > gpurun_out/ncu_full_gemm.log 2>&1
echo "gemm full capture exit $?"
