#!/bin/bash
# compute-sanitizer passes over the GPU tests (small / medium graphs); output under gpurun_out/
K='cfg1_rgbd or slam_dual or random2_ba or tiny_all_fixed or davis or mid_graph or se3_ops or reproject or fused_update or host_buffer or band_solver or every_band_solver or capacity_plan or device_factor_graph or trajectory'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -k "$K" > gpurun_out/san_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/san_memcheck.log | tail -3
# racecheck: once on the verification build of the band solver (tools/racecheck_verify.sh build; -DBA_VERIFY_SYNC: every
# thread behind one of the back substitution's mbarriers arrives on it itself — the form racecheck can follow), once on
# the product build (one lane arrives after a __syncwarp: racecheck reports those producer / consumer pairs)
K2='cfg1_rgbd or slam_dual or tiny_bounds or random_rgbd or mid_graph or every_band_solver or davis or band_solver_failure'
for lib in libbatrack_ba_verify.so libbatrack_ba.so; do
  BATRACK_B200_LIB=$PWD/batrack_b200/$lib timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis \
      python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$K2" > gpurun_out/san_racecheck_$lib.log 2>&1
  echo "racecheck $lib rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/san_racecheck_$lib.log | tail -3
done
