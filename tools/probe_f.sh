#!/bin/bash
timeout 300 python tools/probe.py pca 2>&1 | grep -E "subspace|cov |FAIL|PASS|EXCEPTION"
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "pca" 2>&1 | tail -4
