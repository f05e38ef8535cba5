#!/bin/bash
# ncu --set full captures of the trunk kernels the round-1 verdict names (run under gpurun, one GPU).
cd /root/repo
mkdir -p gpurun_out
python tools/trunk_once.py 256 3 > gpurun_out/r02_trunk_plain.log 2>&1 || exit 1
NCU="ncu --set full --import-source on --clock-control none"
# chain kernels of the second call: L1.0 (shortcut folded), L1.1, L1.2 (<128>), L2.0
$NCU -k regex:conv_chain_kernel -s 7 -c 4 -f -o gpurun_out/r02_chain python tools/trunk_once.py 256 2 > gpurun_out/r02_ncu_chain.log 2>&1
# conv_gemm2 launches of the second call: index 3 = L2.1 conv2 3x3 (N=128), index 8 = L3.0 conv3+res, 9 = L3.1 conv1
$NCU -k regex:conv_gemm2_kernel -s 37 -c 1 -f -o gpurun_out/r02_l2c2 python tools/trunk_once.py 256 2 > gpurun_out/r02_ncu_l2c2.log 2>&1
$NCU -k regex:conv_gemm2_kernel -s 42 -c 2 -f -o gpurun_out/r02_l3c3 python tools/trunk_once.py 256 2 > gpurun_out/r02_ncu_l3c3.log 2>&1
$NCU -k regex:stem_pool -s 1 -c 1 -f -o gpurun_out/r02_stem python tools/trunk_once.py 256 2 > gpurun_out/r02_ncu_stem.log 2>&1
ls -la gpurun_out/*.ncu-rep
