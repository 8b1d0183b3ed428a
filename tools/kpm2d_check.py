"""Developer check: two-kernel 2-D KPM path (kpm2d.cu) against the single-kernel path (kpm.cu) and timing of both."""
import os, sys
import ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fk_mc_b200 as fk

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
rng = np.random.default_rng(1)
for kind, L, beta, U in [("cubic2d", 32, 20.0, 2.0), ("cubic2d", 16, 10.0, 2.0), ("cubic2d", 24, 10.0, 2.0), ("triangular", 24, 10.0, 2.0),
                         ("honeycomb", 24, 10.0, 2.0), ("honeycomb", 16, 5.0, 4.0), ("triangular", 32, 5.0, 1.0)]:
    c = fk.Context(kind, L, max_batch=B)
    N = c.N
    M, G = fk.cheb_sizes(N, 2.2)
    f = (rng.random((B, N)) < 0.5).astype(np.int32)
    f[0] = 0
    f[1] = 1
    res = {}
    for v1 in (1, 0):
        c.set_option("kpm_v1", v1)
        r = c.logz_kpm(f, U, U / 2, beta, M, G)
        c.profile_enable(True); c.profile_reset()
        for _ in range(3):
            r = c.logz_kpm(f, U, U / 2, beta, M, G)
        t = {k: c.profile_get(k) for k in ("kpm", "kpm_lanczos", "kpm_moments")}
        c.profile_enable(False)
        st = np.zeros(B, np.int32)
        c.lib.fkmc_kpm_last_steps(c.h, B, st.ctypes.data_as(C.POINTER(C.c_int32)))
        res[v1] = r
        print("%-10s L=%d M=%d %s: kpm %.3f ms  lanczos %.3f  moments %.3f | steps min %d mean %.1f max %d" % (
            kind, L, M, "v1" if v1 else "v2", t["kpm"][0] / max(1, t["kpm"][1]), t["kpm_lanczos"][0] / max(1, t["kpm_lanczos"][1]),
            t["kpm_moments"][0] / max(1, t["kpm_moments"][1]), st.min(), st.mean(), st.max()), flush=True)
    a, b = res[1], res[0]
    sc = np.abs(a["logZ"]).max()
    print("   max |dlogZ|/|logZ| %.2e  moments %.2e  e_min/e_max %.2e" % (
        np.abs(a["logZ"] - b["logZ"]).max() / sc, np.abs(a["moments"] - b["moments"]).max(),
        max(np.abs(a["e_min"] - b["e_min"]).max(), np.abs(a["e_max"] - b["e_max"]).max())), flush=True)
    c.close()
