#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q -k "chain or embeddings_match or per_layer or batch_invariant or model_callable" 2>&1 | tail -8 | tee gpurun_out/ds_check.log
for nf in 1 0; do echo "IRP_NO_DS_FUSE=$nf"; IRP_NO_DS_FUSE=$nf timeout 300 python tools/trunk_once.py 256 5 2>&1 | tail -1; done | tee -a gpurun_out/ds_check.log
