#!/usr/bin/env python
"""Golden fp64 results for the graphs whose oracle run takes minutes (cfg3: 10 chained LM iterations as BASELINE.json
configs[2] states; cfg5: 2 iterations of the 1024-keyframe graph of configs[4]) — produced ONCE in the build
container by the sparse-aware fp64 oracle (oracle/ba_oracle.py mode="sparse"; the reference's own dense E needs
6.4 GB per temporary at cfg5, SURVEY.md §8c "Oracle limits"), which tests/test_oracle.py pins against the reference's
own code on the small fixtures. Poses are stored in fp64, disparities in fp32 (6e-8 relative, four orders below the
1e-4 bar) to keep the files small.

    python tests/golden/make_golden_large.py            # writes tests/golden/cfg3_x10_sparse64.npz, cfg5_x2_sparse64.npz
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import synth          # noqa: E402
from oracle import ba_oracle            # noqa: E402

for name, iters, out in (("cfg3", 10, "cfg3_x10_sparse64.npz"), ("cfg5", 2, "cfg5_x2_sparse64.npz")):
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    t0 = time.time()
    prob = synth.make_config(name)
    P, D = ba_oracle.run_sequence(prob, [prob.weights] * iters, [False] * iters, torch.float64, mode="sparse")
    np.savez_compressed(os.path.join(HERE, out), poses=P, disps=D.astype(np.float32), edges=np.int64(prob.E),
                        seed=np.int64(0), iters=np.int64(iters))
    print(f"{name}: {iters} iterations in {time.time() - t0:.0f} s -> {out}")
