#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_variants.py tests/test_gpu_parity.py -m gpu -x -q -k "u8 or div2 or random_geometry" ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log
timeout 120 python tools/ab_bench.py u8 2>&1 | grep -E "b=0|segs=64" > gpurun_out/ab_u8.txt
cat gpurun_out/ab_u8.txt
