timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python tools/stage_times.py cfg3 2>&1 | tail -2
timeout 120 python tools/stage_times.py davis 2>&1 | tail -2
