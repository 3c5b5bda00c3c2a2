#!/bin/bash
# Experiment builds of ONE source file: scripts/build_exp_libs.sh egnn_tc2 "0 1 2 4" [-DOTHER]  ->  immunostruct_b200/build/exp/lib_<file>_<n>.so
# (the file compiled with -DIS_EXP=<n>, linked against the objects of the production build; load with IS_B200_DEBUG_LIB)
set -e
cd "$(dirname "$0")/.."
F=$1; LIST=$2; EXTRA=$3
OUT=immunostruct_b200/build/exp
mkdir -p $OUT
python -c "from immunostruct_b200 import build; build.build()"
OTHERS=$(ls immunostruct_b200/build/*.o | grep -v "/$F.o")
pids=()
for n in $LIST; do
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DIS_NO_FAST_MATH -DIS_EXP=$n $EXTRA \
       -I immunostruct_b200/csrc -I include -c immunostruct_b200/csrc/$F.cu -o $OUT/${F}_$n.o && \
    nvcc -shared -o $OUT/lib_${F}_$n.so $OUT/${F}_$n.o $OTHERS -gencode arch=compute_100a,code=sm_100a -lcudart ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
ls -la $OUT/*.so
