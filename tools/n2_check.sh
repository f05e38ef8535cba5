#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q -k "wds or preprocess or val_transform or two_pass or classifier or process_image" 2>&1 | tail -15 | tee gpurun_out/n2_check.log
timeout 300 python tools/preprocess_sweep.py 2>/dev/null | tail -3 | tee -a gpurun_out/n2_check.log
