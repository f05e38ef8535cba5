#!/bin/bash
# GPU box: full gpu test-suite, smoke, bench (both arms). Logs land in gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench_ref rc=$?"; tail -c 1500 gpurun_out/bench_ref.log
timeout 900 python tools/bench_classify.py --steps 3 --warmup 3 > gpurun_out/bench_classify.log 2> gpurun_out/bench_classify.err; echo "bench_classify rc=$?"; tail -c 2500 gpurun_out/bench_classify.log; tail -n 5 gpurun_out/bench_classify.err
