#!/bin/bash
mkdir -p gpurun_out
python tools/trunk_once.py 256 3 > gpurun_out/trunk_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv\|avgpool\|stem_ -s 150 -c 50 --csv --log-file gpurun_out/launches_trunk.csv python tools/trunk_once.py 256 3 > gpurun_out/trunk_ncu.log 2>&1
echo "rc=$?"; cat gpurun_out/trunk_plain.log | tail -2; tail -3 gpurun_out/trunk_ncu.log
