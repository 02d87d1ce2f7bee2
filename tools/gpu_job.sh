#!/bin/bash
# Round-2 profiling job (one gpurun call): ncu --set full of the headline kernel at the bench batch (64 frames) and of the
# new row-streaming resize, plus the launch list of a short bench run.  Outputs land in gpurun_out/ and are summarised
# into profiles/ by tools/ncu_summary.py.
mkdir -p gpurun_out
KB_BATCH=64 ncu --set full --clock-control none --import-source on -k regex:fused_rows -s 3 -c 1 -f -o gpurun_out/r2_cfg2 python tools/run_op.py cfg2 3 > gpurun_out/r2_cfg2_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:resize_rows -s 3 -c 1 -f -o gpurun_out/r2_resize_rows python tools/run_op.py resize_f32 3 > gpurun_out/r2_resize_rows_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:warp_bilinear_x4 -s 3 -c 1 -f -o gpurun_out/r2_warp_x4 python tools/run_op.py warp 3 > gpurun_out/r2_warp_x4_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu --quick > gpurun_out/bench_under_ncu_r2.log 2>&1
tail -2 gpurun_out/r2_cfg2_ncu.log
