#!/bin/bash
# Scratch driver for one gpurun call (edited per experiment; the durable scripts are tools/run_op.py and bench.py).
python tools/e2e_sweep.py 2>&1 | tail -14
