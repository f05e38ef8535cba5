#!/bin/bash
export IRP_B200_PARTIAL=1
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu -k "chain" > gpurun_out/pytest_chain.log 2>&1; echo "chain rc=$?"; tail -n 5 gpurun_out/pytest_chain.log
timeout 300 python tools/chain_bench.py 2>&1 | tail -6
