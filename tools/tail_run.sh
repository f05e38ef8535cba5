#!/bin/bash
cd /root/repo
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/tail_timeline.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -24
bash tools/run_ngpu.sh $N | tail -c 700
