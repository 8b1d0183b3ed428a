"""One launch of each hot kernel of the c5 workload, for ncu captures: python tools/prof_c5.py [B_kpm] [B_ed]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fk_mc_b200 as fk
Bk = int(sys.argv[1]) if len(sys.argv) > 1 else 592
Be = int(sys.argv[2]) if len(sys.argv) > 2 else 296
c = fk.Context("cubic2d", 32, max_batch=max(Bk, Be))
rng = np.random.default_rng(0)
f = (rng.random((max(Bk, Be), c.N)) < 0.5).astype(np.int32)
M, G = fk.cheb_sizes(c.N)
for _ in range(2):
    k = c.logz_kpm(f[:Bk], 2.0, 1.0, 20.0, M, G)
    r = c.logz_ed(f[:Be], 2.0, 1.0, 20.0)
print(k["logZ"][:2], r["logZ"][:2])
