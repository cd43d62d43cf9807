timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
BA_STREAM=0 BA_TRACE=20 timeout 120 python tools/stage_times.py cfg3 2>&1 | grep -E "BA_TRACE\] side|pose\+depth|rror"
timeout 120 python tools/stage_times.py cfg3 2>&1 | grep -E "pose\+depth|rror"; timeout 120 python tools/stage_times.py davis 2>&1 | grep -E "pose\+depth|rror"
BA_SOLVE_NW=12 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mid_graph or band_solver or headline" 2>&1 | tail -2
