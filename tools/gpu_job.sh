#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_variants.py tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/ab_bench.py u8 blur > gpurun_out/ab_bench_u8.txt 2>&1
cat gpurun_out/ab_bench_u8.txt
NCU="ncu --set full --clock-control none --import-source on -s 3 -c 1 -f"
timeout 300 $NCU -k regex:blur_u8_stream -o gpurun_out/r2_blur_u8_stream python tools/run_op.py blur_u8 3 > gpurun_out/r2_blur_u8_ncu.log 2>&1
