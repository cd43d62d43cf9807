for iso in 0 1; do
  echo "== ISO $iso"
  BA_SOLVE_ISO=$iso BA_TRACE=20 timeout 120 python tools/stage_times.py cfg3 2>&1 | grep -E "BA_TRACE|pose\+depth"
done
BA_SOLVE_ISO=1 timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
