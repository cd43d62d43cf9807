timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python tools/stage_times.py cfg3 2>&1 | grep -E "pose\+depth|rror"; timeout 120 python tools/stage_times.py davis 2>&1 | grep -E "pose\+depth|rror"
# serialised execution (every launch blocking): the stand-by launch must take over and give the right answer
CUDA_LAUNCH_BLOCKING=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "headline or mid_graph or davis or 1024" 2>&1 | tail -3
timeout 200 python bench.py --no-cpu-baseline --steps 100 2>/dev/null | python tools/sumbench.py
