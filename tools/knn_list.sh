#!/bin/bash
mkdir -p gpurun_out
for mode in 0 1; do
IRP_KNN_EXHAUSTIVE=$mode ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__inst_executed_op_shared_atom.sum --clock-control none -k regex:knn_ -s 4 -c 2 --csv --log-file gpurun_out/knn_list_$mode.csv python tools/pca_once.py > gpurun_out/pca_ncu.log 2>&1
echo "mode $mode rc=$?"; grep -v "^==" gpurun_out/knn_list_$mode.csv | cut -d, -f5,9,13,15
done
