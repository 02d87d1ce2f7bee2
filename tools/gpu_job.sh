#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_variants.py tests/test_gpu_parity.py -m gpu -x -q -k "blur or u8" ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log
timeout 200 python tools/ab_bench.py blur > gpurun_out/ab_blur.txt 2>&1
cat gpurun_out/ab_blur.txt
