import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fk_mc_b200 as fk
c = fk.Context("cubic2d", 32, max_batch=8)
rng = np.random.default_rng(0)
f = (rng.random((8, c.N)) < 0.5).astype(np.int32)
r = c.logz_ed(f, 2.0, 1.0, 20.0, want_caches=True)
x = r["cached_fermi"].astype(np.int64)
it, nn = x // 1000, x % 1000
print("its: mean %.1f max %d; newton evals: mean %.1f max %d" % (it.mean(), it.max(), nn.mean(), nn.max()))
w = it.reshape(8, 32, 32)
print("per-warp max its: mean %.1f" % w.max(axis=2).mean(), " hist of its:", np.bincount(it.ravel())[:60])
print("newton hist:", np.bincount(nn.ravel())[:40])
