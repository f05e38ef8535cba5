#!/bin/bash
python tools/pre_once.py 2>&1 | tail -1
IRP_PRE_BANDS=1 python tools/pre_once.py 2>&1 | tail -1
timeout 300 python -m pytest tests -x -q -m gpu -k "preprocess or process_image_directory or whole_stage" 2>&1 | tail -3
python tools/preprocess_sweep.py 2>/dev/null | grep nhwc4p | tail -4
