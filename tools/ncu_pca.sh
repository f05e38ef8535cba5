#!/bin/bash
mkdir -p gpurun_out
python tools/pca_once.py > gpurun_out/pca_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"tridiag|bisect|inverse|mgs|back_transform|assemble|cov_gemm|split_transpose|project|knn|lrd|lof_score|percentile|flag|group_|sqnorm|iota" -s 4200 -c 2100 --csv --log-file gpurun_out/launches_pca.csv python tools/pca_once.py > gpurun_out/pca_ncu.log 2>&1
echo rc=$?; cat gpurun_out/pca_plain.log
