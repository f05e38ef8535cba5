#!/bin/bash
cd /root/repo
bash tools/run_suite.sh
bash tools/ncu_bench_launches.sh
