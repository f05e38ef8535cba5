#!/bin/bash
export IRP_B200_PARTIAL=1
mkdir -p gpurun_out
timeout 240 python tools/probe.py conv_3x3 > gpurun_out/probe3_conv_3x3.log 2>&1; echo "conv_3x3 rc=$?"; tail -n 9 gpurun_out/probe3_conv_3x3.log
timeout 120 python tools/one_conv.py 256 56 64 64 3 1 0 6 2>&1 | tail -1
IRP_NO_PATCH64=1 timeout 120 python tools/one_conv.py 256 56 64 64 3 1 0 6 2>&1 | tail -1
timeout 300 python tools/trunk_once.py 256 5 2>&1 | tail -2
