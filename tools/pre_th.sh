#!/bin/bash
python tools/pre_once.py 2>&1 | tail -1
timeout 300 python -m pytest tests -x -q -m gpu -k "preprocess or process_image_directory" 2>&1 | tail -2
