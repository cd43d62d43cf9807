timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for kp in 1 2 4; do echo "KP=$kp"; BA_EDGE2_KP=$kp timeout 120 python tools/stage_times.py cfg3 2>&1 | tail -2; done
for kp in 2 4 8; do echo "KP=$kp"; BA_EDGE2_KP=$kp timeout 120 python tools/stage_times.py davis 2>&1 | tail -2; done
