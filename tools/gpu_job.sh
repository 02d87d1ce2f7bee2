#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_variants.py -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/ab_bench.py warp > gpurun_out/ab_bench.txt 2>&1
grep -E "remap|warp_perspective|rot3 " gpurun_out/ab_bench.txt
