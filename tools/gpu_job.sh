#!/bin/bash
# Scratch driver for one gpurun call (edited per experiment; the durable scripts are tools/run_op.py and bench.py).
python -m pytest tests -m gpu -x -q -k "video_encode or nv12 or yuyv" 2>&1 | tail -3
