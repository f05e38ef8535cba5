#!/bin/bash
# ncu full-set capture (with source-level sampling) of single conv layers, v2 and v1.
mkdir -p gpurun_out
A="256 56 64 256 1 1 0"
python tools/one_conv.py $A > gpurun_out/one_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 2 -c 1 -f -o gpurun_out/prof_ds_v2 python tools/one_conv.py $A > gpurun_out/one_ncu_v2.log 2>&1
echo "v2 rc=$?"; cat gpurun_out/one_plain.log | tail -1
IRP_CONV_V1=1 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 2 -c 1 -f -o gpurun_out/prof_ds_v1 python tools/one_conv.py $A > gpurun_out/one_ncu_v1.log 2>&1
echo "v1 rc=$?"
A="256 28 128 128 3 1 0"
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 2 -c 1 -f -o gpurun_out/prof_l2c2_v2 python tools/one_conv.py $A > gpurun_out/one_ncu_l2c2_v2.log 2>&1
echo "l2c2 v2 rc=$?"
ls -la gpurun_out/*.ncu-rep
