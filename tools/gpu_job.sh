#!/bin/bash
# One gpurun call: GPU parity suite, A/B timings of the store paths, ncu --set full of the lean warp and the u8 warp.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/ab_bench.py warp u8 > gpurun_out/ab_bench.txt 2>&1
cat gpurun_out/ab_bench.txt
NCU="ncu --set full --clock-control none --import-source on -s 3 -c 1 -f"
timeout 300 $NCU -k regex:warp_bilinear_lean -o gpurun_out/r2_warp_lean python tools/run_op.py warp 3 > gpurun_out/r2_warp_lean_ncu.log 2>&1
timeout 300 $NCU -k regex:warp_perspective_u8 -o gpurun_out/r2_warp_u8 python tools/run_op.py warp_u8 3 > gpurun_out/r2_warp_u8_ncu.log 2>&1
tail -2 gpurun_out/r2_warp_u8_ncu.log
