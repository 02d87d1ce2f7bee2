#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_variants.py -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log
timeout 300 python tools/ab_bench.py warp > gpurun_out/ab_bench.txt 2>&1
grep -E "rot|c=2|warp_perspective a=0" gpurun_out/ab_bench.txt
