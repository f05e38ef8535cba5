#!/bin/bash
export IRP_B200_PARTIAL=1
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu -k "chain" > gpurun_out/pytest_chain.log 2>&1; echo "chain rc=$?"; tail -n 12 gpurun_out/pytest_chain.log
timeout 600 python -m pytest tests -x -q -m gpu -k "trunk or embedding or model_callable" > gpurun_out/pytest_trunk.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest_trunk.log
timeout 300 python tools/trunk_once.py 256 8 2>&1 | tail -1
IRP_CHAIN=0 timeout 300 python tools/trunk_once.py 256 8 2>&1 | tail -1
IRP_CHAIN=2 timeout 300 python tools/trunk_once.py 256 8 2>&1 | tail -1
