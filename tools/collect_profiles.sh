#!/bin/bash
# Round profile collection on the GPU box (run through gpurun from the repo root). Everything lands in gpurun_out/;
# tools/ncu_summary.py + a copy step (run in the build container) turn it into profiles/rNN_*.
set -x
R=${1:-r01}
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
python tools/sumbench.py < gpurun_out/${R}_bench.json
# launch list of the bench command (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --profile --steps 6 --warmup 3 > gpurun_out/${R}_launches.log 2>&1
# one full capture per kernel of the step
ncu --set full --import-source on --clock-control none --kernel-name regex:'k_edge_pass_v2|k_schur|k_solve_band_mma|k_backsub' \
    --launch-skip 8 --launch-count 4 -f -o gpurun_out/${R}_kernels python bench.py --profile --steps 6 --warmup 3 > gpurun_out/${R}_ncu.log 2>&1
python tools/plan_build_time.py davis cfg3 > gpurun_out/${R}_plan_build.txt 2>&1
python tools/stage_times.py davis > gpurun_out/${R}_stage_davis.txt 2>&1
python tools/stage_times.py cfg3 > gpurun_out/${R}_stage_cfg3.txt 2>&1
ls -la gpurun_out/
