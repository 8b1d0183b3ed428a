import sys
sys.path.insert(0, ".")
import numpy as np, fk_mc_b200 as fk
for kind, L, B in (("cubic2d", 32, 1024), ("cubic2d", 16, 4096), ("cubic2d", 24, 1024), ("triangular", 24, 1024), ("honeycomb", 24, 1024)):
    c = fk.Context(kind, L, max_batch=B)
    rng = np.random.default_rng(0)
    f = (rng.random((B, c.N)) < 0.5).astype(np.int32)
    M, G = fk.cheb_sizes(c.N, 2.2)
    c.logz_kpm(f, 2.0, 1.0, 10.0, M, G)
    c.profile_enable(True); c.profile_reset()
    for _ in range(3):
        c.logz_kpm(f, 2.0, 1.0, 10.0, M, G)
    st = c.kpm_last_steps(B)
    print("%s L=%d: lanczos %.3f ms per %d proposals; steps mean %.1f min %d max %d" % (kind, L, c.profile_get("kpm_lanczos")[0] / 3, B, st.mean(), st.min(), st.max()), flush=True)
    c.close()
