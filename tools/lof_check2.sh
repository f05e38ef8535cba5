#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "lof or detect or whole_stage or centroid" > gpurun_out/pytest_lof.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest_lof.log
python tools/pca_once.py 2>&1 | tail -1
IRP_KNN_EXHAUSTIVE=1 python tools/pca_once.py 2>&1 | tail -1
python - <<'PY'
import os, sys
sys.path.insert(0, "image-recognition-pipeline_b200"); sys.path.insert(0, ".")
import numpy as np, torch
from irp_b200 import ops
from oracle import synth
# bitwise equality of the filtered search with the exhaustive fp64 search, incl. exact duplicates
z, y = synth.clustered_points(20000, 50, 10, seed=5)
z[:300] = z[1000:1300]          # exact duplicates -> ties
zt = torch.from_numpy(z).cuda(); ids = torch.from_numpy(y.astype(np.int32)).cuda()
res = {}
for mode in ("0", "1"):
    os.environ["IRP_KNN_EXHAUSTIVE"] = mode
PY
