#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python tools/trunk_once.py 256 3 > gpurun_out/trunk_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:conv\|avgpool\|stem_ -s 148 -c 49 --csv --log-file gpurun_out/trunk_traffic_v6.csv python tools/trunk_once.py 256 3 > gpurun_out/trunk_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/trunk_plain.log
