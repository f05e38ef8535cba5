#!/bin/bash
cd /root/repo
for blocks in 3 7; do
for mb in 256 128 64 32 16; do
  echo "blocks=$blocks micro=$mb"
  IRP_TRUNK_BLOCKS=$blocks IRP_MICRO_BATCH=$mb timeout 300 python tools/trunk_once.py 256 5 2>&1 | tail -1
done; done | tee gpurun_out/early_sweep.log
