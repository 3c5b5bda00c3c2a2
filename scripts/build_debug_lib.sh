#!/bin/bash
# Debug build of the kernel library with extra nvcc flags (default -DIS_TRACE: timeline stamps in the edge forward kernel)
# into immunostruct_b200/build/libimmunostruct_b200_debug.so; load it with IS_B200_DEBUG_LIB=<that path>.
set -e
cd "$(dirname "$0")/.."
EXTRA=${1:--DIS_TRACE}
OUT=immunostruct_b200/build/debug
mkdir -p $OUT
rm -f $OUT/*.o
pids=()
for f in immunostruct_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DIS_NO_FAST_MATH $EXTRA \
       -I immunostruct_b200/csrc -I include -c $f -o $OUT/$(basename ${f%.cu}).o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -o immunostruct_b200/build/libimmunostruct_b200_debug.so $OUT/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
echo immunostruct_b200/build/libimmunostruct_b200_debug.so
