#!/bin/bash
mkdir -p gpurun_out
python tools/pre_once.py > gpurun_out/pre_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:resample -s 6 -c 6 --csv --log-file gpurun_out/pre_launches.csv python tools/pre_once.py > gpurun_out/pre_ncu.log 2>&1
echo rc=$?; grep -v "^==" gpurun_out/pre_launches.csv | cut -d, -f5,9,15 | tail -7
python - <<'PY'
import sys, os
sys.path.insert(0, "image-recognition-pipeline_b200"); sys.path.insert(0, ".")
import numpy as np
from oracle import synth
for n in (256, 27000):
    hw = np.minimum(synth.mixed_resolution_sizes(n, seed=0), 1200)
    short = hw.min(1)
    print(n, "images; short side > 580 (generic path):", int((short > 580).sum()), "max", hw.max(0), "mean", hw.mean(0))
PY
