#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q -k "val_transform or classifier or preprocess" 2>&1 | tail -15 | tee gpurun_out/n1_check.log
