"""Per-launch times of the two KPM kernels at the c5 shape through the local seam call (FKMC_LIB selects a developer build):
python tools/kpm_time.py [B]"""
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
c.profile_enable(True)
c.profile_reset()
chk = 0.0
for it in range(12):
    f2 = f.copy()
    f2[np.arange(B), rng.integers(0, c.N, size=B)] ^= 1
    r2 = c.logz_kpm_local(f2, 2.0, 1.0, 20.0, M, G, f_ref=f, state_ref=r["state"])
    chk += float(r2["logZ"].sum())
out = {k: c.profile_get(k) for k in ("kpm_lanczos", "kpm_moments")}
print({k: round(v[0] / max(v[1], 1), 4) for k, v in out.items()}, "logZ checksum %.10e" % chk)
