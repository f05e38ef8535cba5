#!/bin/bash
export IRP_B200_PARTIAL=1
for t in conv_flat conv_3x3 conv_s2 stem; do
  timeout 240 python tools/probe.py $t 2>&1 | grep -E "FAIL|PASS|EXCEPTION|bad " | head -8
done
timeout 400 python tools/probe.py resnet 2>&1 | grep -E "embed cos|batch 256|FAIL|EXCEPTION|Error" | head
for mb in 256 128 64 32; do
  echo "== micro-batch $mb"; IRP_MICRO_BATCH=$mb timeout 200 python tools/trunk_once.py 256 5 2>&1 | tail -1
done
