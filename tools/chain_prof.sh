#!/bin/bash
mkdir -p gpurun_out
python tools/chain_bench.py > gpurun_out/chain_bench.log 2>&1; echo "rc=$?"; cat gpurun_out/chain_bench.log
python tools/chain_bench.py L1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_chain -s 2 -c 1 -f -o gpurun_out/prof_chain python tools/chain_bench.py L1 > gpurun_out/chain_ncu.log 2>&1
echo "ncu rc=$?"
