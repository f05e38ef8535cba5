#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q -k "pca or whole_stage or stage" 2>&1 | tail -15 | tee gpurun_out/pca_ab.log
