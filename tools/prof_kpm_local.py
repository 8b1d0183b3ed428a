"""One full and two local KPM evaluations at the c5 shape, for ncu captures of lanczos2d_kernel / kpm_moments2d_kernel:
python tools/prof_kpm_local.py [B]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fk_mc_b200 as fk
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
c = fk.Context("cubic2d", 32, max_batch=B)
rng = np.random.default_rng(0)
f = (rng.random((B, c.N)) < 0.5).astype(np.int32)
M, G = fk.cheb_sizes(c.N)
r = c.logz_kpm_local(f, 2.0, 1.0, 20.0, M, G)
for _ in range(2):
    f2 = f.copy()
    f2[np.arange(B), rng.integers(0, c.N, size=B)] ^= 1
    r2 = c.logz_kpm_local(f2, 2.0, 1.0, 20.0, M, G, f_ref=f, state_ref=r["state"])
print(r["logZ"][:2], r2["logZ"][:2])
