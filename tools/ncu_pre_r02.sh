#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python tools/pre_once.py 256 > gpurun_out/r02_pre_plain.log 2>&1 || { tail -5 gpurun_out/r02_pre_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:resample\|fused\|plan -s 6 -c 9 --csv --log-file gpurun_out/r02_pre_launches.csv python tools/pre_once.py 256 > gpurun_out/r02_pre_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:resample_fused -s 3 -c 1 -f -o gpurun_out/r02_pre_fused python tools/pre_once.py 256 > gpurun_out/r02_pre_ncu2.log 2>&1
cat gpurun_out/r02_pre_plain.log
