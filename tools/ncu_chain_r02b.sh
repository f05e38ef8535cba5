#!/bin/bash
# ncu --set full capture of the layer1 junction kernels (second trunk call): L1.0 (shortcut folded), L1.1
cd /root/repo; mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none"
$NCU -k regex:conv_chain_kernel -s 7 -c 2 -f -o gpurun_out/r02b_chain python tools/trunk_once.py 256 2 > gpurun_out/r02b_ncu_chain.log 2>&1
ls -la gpurun_out/*.ncu-rep
