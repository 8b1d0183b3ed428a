import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, fk_mc_b200 as fk
for L, B in ((16, 4096), (24, 1024)):
    c = fk.Context("cubic2d", L, max_batch=B)
    rng = np.random.default_rng(0)
    f = (rng.random((B, c.N)) < 0.5).astype(np.int32)
    for w in (0, 2, 3, 4, 6, 8):
        c.set_option("sb2st_warps", w)
        c.logz_ed(f, 2.0, 1.0, 10.0)
        c.profile_enable(True); c.profile_reset()
        for _ in range(3):
            c.logz_ed(f, 2.0, 1.0, 10.0)
        ms, n = c.profile_get("sb2st")
        c.profile_enable(False)
        print("L=%d warps=%d sb2st %.3f ms" % (L, w, ms / n), flush=True)
    c.close()
