#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "lof or detect or whole_stage or centroid" > gpurun_out/pytest_lof.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest_lof.log
python tools/pca_once.py 2>&1 | tail -1
IRP_KNN_GENERIC=1 python tools/pca_once.py 2>&1 | tail -1
