#!/bin/bash
mkdir -p gpurun_out
python tools/pca_once.py > gpurun_out/pca_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:knn_filter -s 5 -c 1 -f -o gpurun_out/prof_knn2 python tools/pca_once.py > gpurun_out/pca_ncu.log 2>&1
echo "ncu rc=$?"
