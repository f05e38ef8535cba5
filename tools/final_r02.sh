#!/bin/bash
# Round-2 evidence pass on one B200: tests (twice), bench (both arms), sweeps, ncu launch lists.
cd /root/repo; mkdir -p gpurun_out
for i in 1 2; do timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_pytest_final_$i.log; cat gpurun_out/r02_pytest_final_$i.log; done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 400 gpurun_out/r02_bench_final.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -c 300 gpurun_out/r02_bench_reference.json
python tools/preprocess_sweep.py > gpurun_out/r02_preprocess_sweep.jsonl 2> gpurun_out/r02_preprocess_sweep.err; tail -2 gpurun_out/r02_preprocess_sweep.jsonl
python tools/wds_sweep.py > gpurun_out/r02_wds_lanczos_sweep.jsonl 2> gpurun_out/r02_wds_sweep.err; tail -2 gpurun_out/r02_wds_lanczos_sweep.jsonl
python tools/bench_classify.py > gpurun_out/r02_bench_classify.json 2> gpurun_out/r02_bench_classify.err; tail -c 300 gpurun_out/r02_bench_classify.json
bash tools/ncu_launches.sh r02_ncu_trunk_traffic
bash tools/ncu_pre_r02.sh
bash tools/ncu_bench_launches.sh
