#!/bin/bash
# Scratch driver for one gpurun call (edited per experiment; the durable scripts are tools/run_op.py and bench.py).
python -m pytest tests -m gpu -x -q -k "gauss or sobel or separable or filter" 2>&1 | tail -1
python tools/run_op.py gauss 30; python tools/run_op.py sobel 30
