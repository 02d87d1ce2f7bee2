#!/bin/bash
# Scratch driver for one gpurun call (edited per experiment; the durable scripts are tools/run_op.py and bench.py).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "warp and not u8" 2>&1 | tail -2
for op in warp affine; do
  echo -n "gather-x4 "; KB200_WARP_IMPL=1 python tools/run_op.py $op 30
  echo -n "old-tiled "; KB200_WARP_IMPL=2 python tools/run_op.py $op 30
  echo -n "tiled-x4  "; python tools/run_op.py $op 30
done
