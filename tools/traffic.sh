#!/bin/bash
mkdir -p gpurun_out
python tools/preprocess_sweep.py > gpurun_out/preprocess_sweep.jsonl 2> gpurun_out/preprocess_sweep.err; echo "sweep rc=$?"; tail -4 gpurun_out/preprocess_sweep.jsonl
python tools/trunk_once.py 256 3 > gpurun_out/trunk_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:conv\|avgpool\|stem_ -s 150 -c 50 --csv --log-file gpurun_out/trunk_traffic.csv python tools/trunk_once.py 256 3 > gpurun_out/trunk_ncu.log 2>&1
echo "ncu rc=$?"
