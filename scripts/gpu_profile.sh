#!/bin/bash
# ncu evidence for the bench command: launch list (per-launch device time) of the timed inference region
# for both arithmetic modes + one full capture of the dominant kernel + the launch list of one training
# step.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
for prec in bf16x3 bf16; do
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches_${prec}.csv \
      python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline --profile --precision ${prec} > gpurun_out/ncu_bench_${prec}.log 2>&1
  echo "launch list ${prec} exit $?"
  ncu --set full --clock-control none --import-source on -k regex:edge_fwd_ws -s 2 -c 1 -f -o gpurun_out/prof_edge_ws_${prec} \
      python scripts/prof_edge.py ${prec} > gpurun_out/ncu_full_${prec}.log 2>&1
  echo "full capture ${prec} exit $?"
done
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_train.csv python scripts/prof_train.py bf16x3 > gpurun_out/ncu_train.log 2>&1
echo "train launch list exit $?"
