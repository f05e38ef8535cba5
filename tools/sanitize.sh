#!/bin/bash
# compute-sanitizer memcheck over the convolution kernels at small shapes (one tool per call).
mkdir -p gpurun_out
export IRP_B200_PARTIAL=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -x -q -m gpu \
  -k "conv1x1_chain and (128-64 or 300-64 or 1000-64) or conv2d_matches_torch and (1-16-64 or 3-7-128 or 2-56-64-64-3 or 2-56-128)" \
  > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -n 12 gpurun_out/sanitizer_memcheck.log
