#!/bin/bash
export IRP_B200_PARTIAL=1
timeout 300 python tools/probe.py preprocess 2>&1 | grep -E "mismatch|preprocess\]|PASS|FAIL|EXC"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "preprocess" 2>&1 | tail -4
