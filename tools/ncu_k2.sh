timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 120 python tools/stage_times.py cfg3 2>&1 | tail -2
timeout 120 python tools/stage_times.py davis 2>&1 | tail -2
ncu --set full --import-source on --clock-control none --kernel-name regex:k_schur --launch-skip 6 --launch-count 1 -f -o gpurun_out/k2_cfg3 python tools/stage_times.py cfg3 > gpurun_out/ncu_k2_cfg3.log 2>&1
