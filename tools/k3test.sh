for ge in 24 40 64 1000; do echo "GEND=$ge"; BA_STREAM_GEND=$ge timeout 120 python tools/stage_times.py cfg3 2>&1 | grep -E "pose\+depth|rror"; done
BA_STREAM_GEND=40 timeout 120 python tools/stage_times.py davis 2>&1 | grep -E "pose\+depth|rror"
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
