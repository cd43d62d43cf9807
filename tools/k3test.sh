for nw in 8 12; do
echo "== NW $nw"
BA_SOLVE_NW=$nw BA_TRACE=20 timeout 120 python tools/stage_times.py cfg3 2>&1 | grep -E "BA_TRACE|pose\+depth|rror"
BA_SOLVE_NW=$nw timeout 120 python tools/stage_times.py davis 2>&1 | grep -E "pose\+depth|rror"
done
BA_SOLVE_NW=12 BA_SOLVE_ISO=0 timeout 120 python tools/stage_times.py cfg3 2>&1 | grep -E "pose\+depth|rror"
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
