#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
N=$1
for extra in "" "" ""; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 12 --warmup 3 $extra > gpurun_out/stall_$N.json 2> gpurun_out/stall_$N.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/stall_$N.json').read().strip().splitlines()[-1])
print("extra='$extra'", d['value'], d['stage_ms']['per_step'])
PY
done
