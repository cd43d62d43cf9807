#!/bin/bash
# racecheck on a verification build of the band solver (-DBA_VERIFY_SYNC: every thread behind one of the back
# substitution's mbarriers arrives on it itself), next to the product build, on the tests that run that solver.
# Build here (no GPU needed): bash tools/racecheck_verify.sh build ; run on the GPU box: bash tools/racecheck_verify.sh
set -e
cd "$(dirname "$0")/.."
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -cudart static"
if [ "$1" = build ]; then
  $NVCC $FLAGS -DBA_VERIFY_SYNC -c batrack_b200/csrc/ba_solve_diag.cu -o batrack_b200/csrc/ba_solve_diag_verify.o
  OBJS=$(ls batrack_b200/csrc/*.o | grep -v ba_solve_diag)
  $NVCC -shared -cudart static -o batrack_b200/libbatrack_ba_verify.so $OBJS batrack_b200/csrc/ba_solve_diag_verify.o
  exit 0
fi
K='every_band_solver or mid_graph or band_solver_failure'
for lib in libbatrack_ba_verify.so libbatrack_ba.so; do
  echo "## $lib"
  BATRACK_B200_LIB=$PWD/batrack_b200/$lib timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis \
      python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$K" > gpurun_out/racecheck_$lib.log 2>&1 || true
  grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/racecheck_$lib.log | tail -3
done
