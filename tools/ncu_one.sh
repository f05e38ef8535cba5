#!/bin/bash
# usage: ncu_one.sh <kernel regex> <skip> <count> <out name>   -- ncu --set full of trunk kernels at batch 256
cd /root/repo; mkdir -p gpurun_out
python tools/trunk_once.py 256 3 > gpurun_out/$4_plain.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:$1 -s $2 -c $3 -f -o gpurun_out/$4 python tools/trunk_once.py 256 2 > gpurun_out/$4_ncu.log 2>&1
cat gpurun_out/$4_plain.log
