#!/bin/bash
# usage: run_cfgs.sh N "cfg ..." [steps] [warmup]   -- bench.py under torchrun on N GPUs for each named config
cd /root/repo
mkdir -p gpurun_out
N=$1; CFGS=$2; STEPS=${3:-5}; WARM=${4:-3}
for c in $CFGS; do
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps $STEPS --warmup $WARM --config $c --no-cpu-baseline \
      > gpurun_out/r02_${c}_n${N}.json 2> gpurun_out/r02_${c}_n${N}.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps $STEPS --warmup $WARM --config $c > gpurun_out/r02_${c}_n${N}.json 2> gpurun_out/r02_${c}_n${N}.err
  fi
  echo "== $c N=$N rc=$?"; tail -c 900 gpurun_out/r02_${c}_n${N}.json; tail -2 gpurun_out/r02_${c}_n${N}.err
done
