#!/bin/bash
mkdir -p gpurun_out
python tools/l1_bench.py 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:l1_block -s 2 -c 1 -f -o gpurun_out/prof_l1 python tools/l1_bench.py > gpurun_out/l1_ncu.log 2>&1
echo "ncu rc=$?"
