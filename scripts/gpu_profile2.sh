#!/bin/bash
# Round-2 ncu evidence: launch lists of one training step and one cancer fine-tune step, full captures of the
# kernels under work (edge backward, node kernel, TMA GEMM).  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
python scripts/prof_train.py bf16x3 train fused > gpurun_out/wall_train.log 2>&1; tail -2 gpurun_out/wall_train.log
python scripts/prof_train.py bf16x3 comparative fused > gpurun_out/wall_cmp.log 2>&1; tail -2 gpurun_out/wall_cmp.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_train.csv python scripts/prof_train.py bf16x3 train fused > gpurun_out/ncu_train.log 2>&1
echo "train launch list exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_cmp.csv python scripts/prof_train.py bf16x3 comparative fused > gpurun_out/ncu_cmp.log 2>&1
echo "comparative launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:edge_bwd_tc -s 8 -c 1 -f -o gpurun_out/prof_edge_bwd \
    python scripts/prof_train.py bf16x3 train fused > gpurun_out/ncu_full_edge_bwd.log 2>&1
echo "edge_bwd full capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:node_post_pre_tc -s 27 -c 1 -f -o gpurun_out/prof_node \
    python scripts/prof_node.py > gpurun_out/ncu_full_node.log 2>&1
echo "node full capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tma_kernel -s 4 -c 2 -f -o gpurun_out/prof_gemm \
    python scripts/prof_train.py bf16x3 train fused > gpurun_out/ncu_full_gemm.log 2>&1
echo "gemm full capture exit $?"
