for pf in 0 1; do for tile in 32 64; do echo "pf $pf tile $tile"; BA_SCHUR_PF=$pf BA_SCHUR_TILE=$tile timeout 120 python tools/stage_times.py cfg3 2>&1 | grep "pose+depth"; done; done
BA_SCHUR_PF=0 timeout 120 python tools/stage_times.py davis 2>&1 | grep "pose+depth"
BA_SCHUR_PF=1 timeout 120 python tools/stage_times.py davis 2>&1 | grep "pose+depth"
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
