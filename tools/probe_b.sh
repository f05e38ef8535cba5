#!/bin/bash
mkdir -p gpurun_out
for t in pca lof; do
  timeout 300 python tools/probe.py $t > gpurun_out/probe_$t.log 2>&1; echo "$t rc=$?"
  tail -n 14 gpurun_out/probe_$t.log
done
