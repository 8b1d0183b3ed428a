import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fk_mc_b200 as fk
B = 1184
c = fk.Context("cubic2d", 32, max_batch=B)
rng = np.random.default_rng(0)
f = (rng.random((B, c.N)) < 0.5).astype(np.int32)
for M, G in [(4, 10), (16, 32)]:
    c.logz_kpm(f, 2.0, 1.0, 20.0, M, G)
    c.profile_enable(True); c.profile_reset()
    for _ in range(3):
        c.logz_kpm(f, 2.0, 1.0, 20.0, M, G)
    ms, n = c.profile_get("kpm")
    import ctypes as C
    st = np.zeros(B, np.int32)
    c.lib.fkmc_kpm_last_steps(c.h, B, st.ctypes.data_as(C.POINTER(C.c_int32)))
    print("M=%d: %.3f ms per launch of %d; lanczos steps min %d mean %.1f max %d" % (M, ms / n, B, st.min(), st.mean(), st.max()))
