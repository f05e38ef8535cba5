#!/bin/bash
mkdir -p gpurun_out
python tools/trunk_once.py 256 3 > gpurun_out/trunk_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 188 -c 2 -o gpurun_out/prof_conv python tools/trunk_once.py 256 3 > gpurun_out/trunk_ncu_full.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/trunk_ncu_full.log; ls -la gpurun_out/*.ncu-rep
