#!/bin/bash
# Scratch driver for one gpurun call (edited per experiment; the durable scripts are tools/run_op.py and bench.py).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "nv12 or yuyv or normalize or color or rgb" 2>&1 | tail -2
python tools/run_op.py rgb_nv12 30
python tools/run_op.py normalize 30
