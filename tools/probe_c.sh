#!/bin/bash
mkdir -p gpurun_out
for c in 32 8 4 2; do
  echo "== chunk $c"
  IRP_COV_CHUNK=$c timeout 300 python tools/probe.py pca 2>&1 | grep -E "scatter|subspace|cov " 
done
