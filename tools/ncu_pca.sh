#!/bin/bash
mkdir -p gpurun_out
python tools/pca_once.py > gpurun_out/pca_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 4300 -c 2200 --csv --log-file gpurun_out/launches_pca.csv python tools/pca_once.py > gpurun_out/pca_ncu.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/pca_plain.log
