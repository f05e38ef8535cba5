#!/bin/bash
cd /root/repo
N=${1:-4}
IRP_STAGE_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/trace_bench.log 2> gpurun_out/trace_bench.err
grep "trace rank" gpurun_out/trace_bench.err | sort | head -80
grep '^{' gpurun_out/trace_bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stage_ms'].get('per_step'), d['e2e']['value'])"
