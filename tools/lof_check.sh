#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "lof or detect or whole_stage" > gpurun_out/pytest_lof.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest_lof.log
python tools/pca_once.py > gpurun_out/pca_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:knn_kernel -s 4 -c 2 -f -o gpurun_out/prof_knn python tools/pca_once.py > gpurun_out/pca_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/pca_plain.log
