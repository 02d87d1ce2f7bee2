#!/bin/bash
# Scratch driver for one gpurun call (edited per experiment; the durable scripts are tools/run_op.py and bench.py).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "warp" 2>&1 | tail -2
for op in warp affine; do
  echo -n "x1 "; KB200_WARP_IMPL=x1 python tools/run_op.py $op 30
  echo -n "x4 "; python tools/run_op.py $op 30
done
