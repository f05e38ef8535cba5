#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
N=${1:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps ${STEPS:-3} --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"
tail -c 1200 gpurun_out/bench_${N}gpu.log; tail -n 3 gpurun_out/bench_${N}gpu.err
