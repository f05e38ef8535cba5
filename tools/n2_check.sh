#!/bin/bash
cd /root/repo
timeout 600 python tools/wds_sweep.py 2>/dev/null | tee gpurun_out/wds_sweep.jsonl
