#!/bin/bash
mkdir -p gpurun_out
python tools/trunk_once.py 256 3 > gpurun_out/trunk_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -k regex:conv_gemm\|maxpool\|avgpool -s 165 -c 55 --csv --log-file gpurun_out/launches_b256_warm.csv python tools/trunk_once.py 256 3 > /dev/null 2>&1
echo "rc=$?"
IRP_MICRO_BATCH=64 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -k regex:conv_gemm\|maxpool\|avgpool -s 660 -c 55 --csv --log-file gpurun_out/launches_mb64_warm.csv python tools/trunk_once.py 256 3 > /dev/null 2>&1
echo "rc=$?"
