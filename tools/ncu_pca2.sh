#!/bin/bash
cd /root/repo
timeout 300 python tools/pca_fit_once.py 5 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'lz_|bisect|inverse|cluster_mgs|assemble' --csv --log-file gpurun_out/pca_launches.csv python tools/pca_fit_once.py 1 > gpurun_out/ncu_pca.log 2>&1
tail -3 gpurun_out/ncu_pca.log
