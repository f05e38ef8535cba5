#!/bin/bash
# ncu launch list of one step of bench.py (run AFTER bench.py has exited 0 without ncu in the same call): every
# kernel of libirp_b200.so with its device time, cold-cache and serialised -- compare SHARES with stage_ms.
# -k restricts the skip / count to the library's kernels (the synthetic-workload generator alone launches ~270 k
# torch kernels before the first step).
cd /root/repo
mkdir -p gpurun_out
SKIP=${SKIP:-6100}     # the fold kernels at load + one warm-up step (about 5.9 k launches per step)
COUNT=${COUNT:-5938}
RE='conv|stem_|avgpool|resample|lz_|cov_gemm|split_transpose|col_sum|trace_kernel|bisect|inverse_iter|cluster_mgs|assemble_cov|proj_gemm|project_split|clip_evals|add_count|knn|lrd_|lof_|group_|flag_kernel|sort_key|gather_sorted|sqnorm|iota|unsort|DeviceRadixSort|fold_bn|add_bias'
timeout 900 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 || exit 1
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:$RE" -s $SKIP -c $COUNT --csv \
    --log-file gpurun_out/r02_ncu_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
echo "ncu rc=$?"; grep -c gpu__time_duration gpurun_out/r02_ncu_launches_bench.csv
