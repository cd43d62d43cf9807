timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python tools/stage_times.py cfg3 2>&1 | grep -E "pose\+depth|rror"; timeout 120 python tools/stage_times.py davis 2>&1 | grep -E "pose\+depth|rror"
timeout 200 python bench.py --no-cpu-baseline --steps 100 2>/dev/null | python tools/sumbench.py
