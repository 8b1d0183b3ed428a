"""Run a few dense evaluations for ncu captures: python tools/prof_ed.py L B [reps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fk_mc_b200 as fk
L, B = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
c = fk.Context("cubic2d", L, max_batch=B)
rng = np.random.default_rng(0)
f = (rng.random((B, c.N)) < 0.5).astype(np.int32)
for _ in range(reps):
    r = c.logz_ed(f, 2.0, 1.0, 10.0)
print(r["logZ"][:2])
