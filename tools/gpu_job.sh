#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/ab_bench.py warp > gpurun_out/ab_bench.txt 2>&1
grep -E "remap|warp_perspective a=0" gpurun_out/ab_bench.txt
NCU="ncu --set full --clock-control none --import-source on -s 3 -c 1 -f"
timeout 300 $NCU -k regex:warp_bilinear_lean -o gpurun_out/r2_warp_lean python tools/run_op.py warp 3 > gpurun_out/r2_warp_lean_ncu.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
tail -c 200 gpurun_out/bench_r2.json
