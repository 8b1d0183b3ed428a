import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, fk_mc_b200 as fk
c = fk.Context("cubic2d", 32, max_batch=148)
rng = np.random.default_rng(0)
f = (rng.random((148, c.N)) < 0.5).astype(np.int32)
for w in (1, 2, 4):
    c.set_option("sb2st_warps", w)
    print("warps", w, flush=True)
    r = c.logz_ed(f, 2.0, 1.0, 10.0)
    c.sync()
