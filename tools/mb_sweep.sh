#!/bin/bash
for mb in 256 128 64 32 16; do
  echo "== micro-batch $mb"; IRP_MICRO_BATCH=$mb timeout 200 python tools/trunk_once.py 256 5 2>&1 | tail -1
done
IRP_MICRO_BATCH=64 timeout 300 python tools/probe.py resnet 2>&1 | grep -E "embed cos|batch 256"
