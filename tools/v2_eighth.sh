#!/bin/bash
export IRP_B200_PARTIAL=1
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu -k "l1_block" > gpurun_out/pytest_l1.log 2>&1; echo "l1 rc=$?"; tail -n 12 gpurun_out/pytest_l1.log
timeout 600 python -m pytest tests -x -q -m gpu -k "trunk or embedding or model_callable" > gpurun_out/pytest_trunk.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest_trunk.log
timeout 300 python tools/trunk_once.py 256 8 2>&1 | tail -1
IRP_L1_FUSE=0 timeout 300 python tools/trunk_once.py 256 8 2>&1 | tail -1
