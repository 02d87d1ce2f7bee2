#!/bin/bash
# Profiling job (one gpurun call): ncu --set full of the kernels DESIGN.md argues with, one launch each (tools/run_op.py).
# Outputs land in gpurun_out/; `python tools/ncu_summary.py name=gpurun_out/name.ncu-rep ...` writes profiles/<name>_ncu.csv (and updates
# profiles/traffic.json for r2_cfg2), `python tools/ncu_top.py <rep>` prints the busiest units and the stall breakdown.
# The round job (tests + bench + reference arm + A/B table + ncu launch list) is tools/gpu_round.sh.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -s 3 -c 1 -f"
KB_BATCH=64 timeout 300 $NCU -k regex:fused_rows -o gpurun_out/r2_cfg2 python tools/run_op.py cfg2 3 > gpurun_out/r2_cfg2_ncu.log 2>&1
timeout 200 $NCU -k regex:warp_bilinear_lean -o gpurun_out/r2_warp_lean python tools/run_op.py warp 3 > gpurun_out/r2_warp_lean_ncu.log 2>&1
timeout 200 $NCU -k regex:warp_bilinear_lean -o gpurun_out/r2_remap_lean python tools/run_op.py remap 3 > gpurun_out/r2_remap_ncu.log 2>&1
timeout 200 $NCU -k regex:warp_tiled -o gpurun_out/r2_warp_tiled python tools/run_op.py affine 3 > gpurun_out/r2_tiled_ncu.log 2>&1
timeout 200 $NCU -k regex:sep_filter_stream2 -o gpurun_out/r2_sobel python tools/run_op.py sobel 3 > gpurun_out/r2_sobel_ncu.log 2>&1
timeout 200 $NCU -k regex:fused_rows -o gpurun_out/r2_fused_general python tools/run_op.py fused_general 3 > gpurun_out/r2_fused_general_ncu.log 2>&1
timeout 200 $NCU -k regex:warp_perspective_u8 -o gpurun_out/r2_warp_u8 python tools/run_op.py warp_u8 3 > gpurun_out/r2_warp_u8_ncu.log 2>&1
timeout 200 $NCU -k regex:blur_u8_stream -o gpurun_out/r2_blur_u8_stream python tools/run_op.py blur_u8 3 > gpurun_out/r2_blur_u8_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
