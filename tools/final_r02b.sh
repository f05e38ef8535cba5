#!/bin/bash
# Round-2 (second half) evidence pass on one B200: tests twice, smoke, bench (both arms), classifier bench, configs[3]
# on one GPU, ncu launch list of one bench step.  Everything lands in gpurun_out/.
cd /root/repo; mkdir -p gpurun_out
for i in 1 2; do timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_pytest_final_$i.log; cat gpurun_out/r02_pytest_final_$i.log; done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_smoke_final.log
timeout 600 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 300 gpurun_out/r02_bench_final.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -c 300 gpurun_out/r02_bench_reference.json
timeout 400 python tools/bench_classify.py > gpurun_out/r02_bench_classify.json 2> gpurun_out/r02_bench_classify.err; tail -c 300 gpurun_out/r02_bench_classify.json
timeout 600 python bench.py --config cfg4 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_cfg4_n1.json 2> gpurun_out/r02_bench_cfg4_n1.err; tail -c 300 gpurun_out/r02_bench_cfg4_n1.json
bash tools/ncu_bench_launches.sh
