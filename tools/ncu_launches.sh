#!/bin/bash
# ncu launch list (duration, DRAM bytes, tensor-pipe activity) of one trunk call at batch 256 -> gpurun_out/$1.csv
cd /root/repo
mkdir -p gpurun_out
OUT=${1:-launches_trunk}
python tools/trunk_once.py 256 3 > gpurun_out/${OUT}_plain.log 2>&1 || { tail -5 gpurun_out/${OUT}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:conv\|avgpool\|stem_ -s 45 -c 45 --csv --log-file gpurun_out/${OUT}.csv python tools/trunk_once.py 256 2 > gpurun_out/${OUT}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${OUT}_plain.log
