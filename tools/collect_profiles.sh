#!/bin/bash
# Round profile collection on the GPU box (run through gpurun from the repo root). Everything lands in gpurun_out/;
# tools/ncu_summary.py + a copy step (run in the build container) turn it into profiles/rNN_*.
set -x
R=${1:-r02}
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
# launch list of the bench command (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --profile --steps 6 --warmup 3 > gpurun_out/${R}_launches.log 2>&1
# one full capture per kernel of the step
ncu --set full --import-source on --clock-control none --kernel-name regex:'k_edge_pass_v2|k_schur_tc|k_solve_band_diag|k_backsub' \
    --launch-skip 8 --launch-count 4 -f -o gpurun_out/${R}_kernels python bench.py --profile --steps 6 --warmup 3 > gpurun_out/${R}_ncu.log 2>&1
# the small-system path (DAVIS-like window): tile solver, SIMT Schur, lane-per-track edge pass with position splits
ncu --set full --import-source on --clock-control none --kernel-name regex:'k_edge_pass_v2|k_schur|k_solve_tiles|k_backsub' \
    --launch-skip 8 --launch-count 4 -f -o gpurun_out/${R}_kernels_davis python tools/prof_step.py davis > gpurun_out/${R}_ncu_davis.log 2>&1
python tools/plan_build_time.py davis sintel cfg3 > gpurun_out/${R}_plan_build.txt 2>&1
for w in davis sintel cfg3; do python tools/stage_times.py $w > gpurun_out/${R}_stage_$w.txt 2>&1; done
python tools/solver_ab.py cfg3 diag > gpurun_out/${R}_solver_cfg3.txt 2>&1
python tools/solver_ab.py davis diag,tiles,window > gpurun_out/${R}_solver_davis.txt 2>&1
python tools/solver_ab.py sintel tiles,window > gpurun_out/${R}_solver_sintel.txt 2>&1
python tools/schur_trace.py cfg3 > gpurun_out/${R}_schur_trace.txt 2>&1
python tools/solver_timeline.py > gpurun_out/${R}_solver_timeline.txt 2>&1
python tools/sass_excerpt.py > gpurun_out/${R}_sass.txt 2>&1
ls -la gpurun_out/ | tail -30
