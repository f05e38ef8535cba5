#!/bin/bash
export IRP_B200_PARTIAL=1
timeout 240 python tools/probe.py stem 2>&1 | grep -E "stem|EXC|rror" | head -5
IRP_STEM_MODE=0 timeout 240 python tools/probe.py stem 2>&1 | grep -E "stem mode" | head -3
timeout 300 python tools/probe.py resnet 2>&1 | grep -E "embed cos|batch 256: |EXC|rror" | head -3
