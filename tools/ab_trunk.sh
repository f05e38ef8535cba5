#!/bin/bash
# Developer tool: A/B of two builds of libirp_b200.so on ONE box (alternating processes): trunk ms per 256 images.
# usage: tools/ab_trunk.sh [rounds]   -- A = image-recognition-pipeline_b200/lib_ab/libirp_base.so, B = the in-tree build
cd /root/repo
R=${1:-3}
for r in $(seq 1 $R); do
  IRP_AB_LIB=image-recognition-pipeline_b200/lib_ab/libirp_base.so IRP_B200_PARTIAL=1 python tools/trunk_once.py 256 8 | sed 's/^/A base: /'
  python tools/trunk_once.py 256 8 | sed 's/^/B new : /'
done
