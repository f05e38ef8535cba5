#!/bin/bash
export IRP_B200_PARTIAL=1
for c in 0 21 12 22; do
  echo "== cluster $c"
  export IRP_CLUSTER=$c
  for t in conv_flat conv_3x3 conv_s2; do timeout 240 python tools/probe.py $t 2>&1 | grep -E "FAIL|== |EXCEPTION" | head -6; done
  timeout 300 python tools/probe.py resnet 2>&1 | grep -E "embed cos|batch 256: |EXCEPTION|rror" | head -3
done
