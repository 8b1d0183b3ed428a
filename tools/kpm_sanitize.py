"""Small KPM evaluations for compute-sanitizer (memcheck / racecheck) runs of the csrc/kpm2d.cu kernels."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fk_mc_b200 as fk
rng = np.random.default_rng(3)
for kind, L, M in [("cubic2d", 16, 12), ("triangular", 16, 10), ("honeycomb", 16, 12), ("cubic2d", 32, 16)]:
    c = fk.Context(kind, L, max_batch=6)
    f = (rng.random((6, c.N)) < 0.5).astype(np.int32)
    r = c.logz_kpm(f, 2.0, 1.0, 10.0, M, 2 * M)
    print(kind, L, r["logZ"][:2])
    c.close()
