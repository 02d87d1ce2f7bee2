#!/bin/bash
# One gpurun call: GPU parity suite, the default bench (N=1), the reference arm, A/B table and the ncu launch list.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
tail -c 300 gpurun_out/bench_r2.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_r2.json 2> gpurun_out/bench_ref_r2.err
timeout 300 python tools/ab_bench.py warp u8 blur > gpurun_out/ab_bench.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu --quick > gpurun_out/bench_under_ncu_r2.log 2>&1
