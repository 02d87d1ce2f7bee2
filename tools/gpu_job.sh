#!/bin/bash
# Scratch driver for one gpurun call (edited per experiment; the durable scripts are tools/run_op.py and bench.py).
python -m pytest tests -m gpu -x -q -k "warp and not u8" 2>&1 | tail -1
for pf in 96 128 160 192; do echo -n "pf=$pf "; KB200_WARP_PF=$pf python tools/run_op.py warp 30; done
