#!/bin/bash
mkdir -p gpurun_out
python tools/pre_once.py > gpurun_out/pre_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:resample_fast -s 2 -c 1 -f -o gpurun_out/prof_pre python tools/pre_once.py > gpurun_out/pre_ncu.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/pre_plain.log
