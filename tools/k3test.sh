timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
BA_TRACE=20 timeout 120 python tools/stage_times.py cfg3 2>&1 | grep -E "BA_TRACE|pose\+depth"
timeout 120 python tools/stage_times.py davis 2>&1 | tail -2
