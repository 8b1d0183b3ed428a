import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import fk_mc_b200.binding as bd
bd.LIB_PATH = os.path.join(ROOT, "tools", "_scratch", "libfkmc_timing.so")
import fk_mc_b200 as fk
c = fk.Context("cubic2d", 32, max_batch=1024)
rng = np.random.default_rng(0)
f = (rng.random((1024, c.N)) < 0.5).astype(np.int32)
for _ in range(2):
    c.logz_kpm(f, 2.0, 1.0, 20.0, 16, 32)
c.sync()
