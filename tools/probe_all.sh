#!/bin/bash
# Runs every probe in its own process (a CUDA fault in one must not poison the next).
export IRP_B200_PARTIAL=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
for t in conv_flat conv_3x3 conv_s2 preprocess stem; do
  timeout 240 python tools/probe.py $t > gpurun_out/probe_$t.log 2>&1; echo "$t rc=$?"
  tail -n 12 gpurun_out/probe_$t.log
done
IRP_STEM_MODE=1 timeout 240 python tools/probe.py stem > gpurun_out/probe_stem_im2col.log 2>&1; echo "stem_im2col rc=$?"; tail -n 5 gpurun_out/probe_stem_im2col.log
timeout 400 python tools/probe.py resnet > gpurun_out/probe_resnet.log 2>&1; echo "resnet rc=$?"; tail -n 24 gpurun_out/probe_resnet.log
