#!/bin/bash
# One GPU-box visit: parity tests, smoke, short bench.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/kernel_errors.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
# new kernels first, under a short wall-clock limit (a hung kernel must not eat the box)
timeout 240 python -m pytest tests/test_round2_gpu.py -m gpu -q -x -p no:cacheprovider -k "split_planes or gemm_planes or linear_tc" > gpurun_out/pytest_new.log 2>&1
echo "pytest-new exit: $?" >> gpurun_out/pytest_new.log
tail -15 gpurun_out/pytest_new.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit: $?" >> gpurun_out/bench.err
tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
if [ "$1" == "sanitize" ]; then
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py --smoke > gpurun_out/sanitizer.log 2>&1
  echo "sanitizer exit: $?" >> gpurun_out/sanitizer.log
  tail -15 gpurun_out/sanitizer.log
fi
