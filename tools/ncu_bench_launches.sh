#!/bin/bash
# ncu launch list of one step of bench.py (run AFTER bench.py has exited 0 without ncu in the same call): every
# kernel of the repo's library with its device time, cold-cache and serialised -- compare SHARES with stage_ms.
cd /root/repo
mkdir -p gpurun_out
SKIP=${SKIP:-6600}     # one warm-up step (about 6.5 k launches of libirp_b200 kernels + torch's own)
COUNT=${COUNT:-6700}
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $COUNT --csv \
    --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
echo "ncu rc=$?"; grep -c gpu__time_duration gpurun_out/launches_bench.csv
