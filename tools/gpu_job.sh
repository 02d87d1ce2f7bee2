#!/bin/bash
# memcheck / racecheck of the kernels touched in the second pass of round 2
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_variants.py -m gpu -x -q -k "lean or interior or u8_stream or random_geometry or remap" > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_variants.py -m gpu -x -q -k "gaussian_blur_u8_stream or warp_perspective_lean" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck.log
