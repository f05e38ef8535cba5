#!/bin/bash
# First run of the CTA-pair conv kernel: correctness probes, then per-layer and whole-trunk timing, v2 vs v1.
export IRP_B200_PARTIAL=1
mkdir -p gpurun_out
for t in conv_flat conv_3x3 conv_s2; do
  timeout 240 python tools/probe.py $t > gpurun_out/probe2_$t.log 2>&1; echo "$t rc=$?"
  tail -n 9 gpurun_out/probe2_$t.log
done
timeout 300 python tools/conv_bench.py 256 5 > gpurun_out/conv_bench_v2.log 2>&1; echo "bench v2 rc=$?"; cat gpurun_out/conv_bench_v2.log | tail -25
IRP_CONV_V1=1 timeout 300 python tools/conv_bench.py 256 5 > gpurun_out/conv_bench_v1.log 2>&1; echo "bench v1 rc=$?"; cat gpurun_out/conv_bench_v1.log | tail -25
timeout 300 python tools/trunk_once.py 256 5 2>&1 | tail -2
IRP_CONV_V1=1 timeout 300 python tools/trunk_once.py 256 5 2>&1 | tail -2
