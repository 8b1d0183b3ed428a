import sys
sys.path.insert(0, '.')
import numpy as np, fk_mc_b200 as fk
for L, B in ((26, 1024), (28, 1024), (30, 1024)):
    c = fk.Context("cubic2d", L, max_batch=B)
    rng = np.random.default_rng(0)
    f = (rng.random((B, c.N)) < 0.5).astype(np.int32)
    c.logz_ed(f, 2.0, 1.0, 10.0)
    c.profile_enable(True); c.profile_reset()
    for _ in range(3):
        c.logz_ed(f, 2.0, 1.0, 10.0)
    ms, n = c.profile_get("sb2st")
    print("L=%d N=%d sb2st %.3f ms per %d" % (L, c.N, ms / n, B), flush=True)
    c.close()
