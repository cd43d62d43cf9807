#!/bin/bash
# compute-sanitizer passes over the GPU tests (small / medium graphs); output under gpurun_out/
K='cfg1_rgbd or slam_dual or random2_ba or tiny_all_fixed or davis or mid_graph or se3_ops or reproject or fused_update or host_buffer or band_solver or every_band_solver or capacity_plan or device_factor_graph or trajectory'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -k "$K" > gpurun_out/san_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/san_memcheck.log | tail -3
K2='cfg1_rgbd or slam_dual or tiny_bounds or random_rgbd or mid_graph or every_band_solver or davis'
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$K2" > gpurun_out/san_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/san_racecheck.log | tail -3
