for cfg in "32 128" "64 128" "64 256" "32 256"; do set -- $cfg; echo "tile $1 tu $2"; BA_SCHUR_TILE=$1 BA_SCHUR_TU=$2 timeout 120 python tools/stage_times.py cfg3 2>&1 | grep "pose+depth"; done
BA_SCHUR_TILE=64 timeout 120 python tools/stage_times.py davis 2>&1 | grep "pose+depth"
BA_SCHUR_TILE=32 timeout 120 python tools/stage_times.py davis 2>&1 | grep "pose+depth"
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
