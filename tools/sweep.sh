#!/bin/bash
cd /root/repo
timeout 600 python tools/trunk_batch_sweep.py 2>&1 | tee gpurun_out/trunk_batch_sweep.log
timeout 900 python -m pytest tests -m gpu -x -q -k "scale_out" 2>&1 | tail -5 | tee -a gpurun_out/trunk_batch_sweep.log
