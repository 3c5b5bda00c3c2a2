#!/bin/bash
# ncu evidence for the bench command: launch list (per-launch device time) + one full capture of the
# dominant kernel.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:"${1:-edge_fwd_kernel}" -s 6 -c 2 -f -o gpurun_out/prof_edge_fwd \
    python bench.py --steps 1 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out/
