"""Developer probe: quality (residual, orthogonality) and device time of the eigenvector path.  python tools/eigvec_probe.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fk_mc_b200 as fk
import oracle_lib as o
for kind, L, B, U, beta in [("cubic2d", 16, 64, 2.0, 10.0), ("cubic3d", 8, 16, 4.0, 5.0), ("triangular", 24, 16, 2.0, 10.0), ("cubic2d", 32, 8, 2.0, 20.0)]:
    c = fk.Context(kind, L, max_batch=B)
    n = c.N
    rng = np.random.default_rng(1)
    f = (rng.random((B, n)) < 0.5).astype(np.int32)
    c.eigh(f[:2], U, U / 2, beta)
    c.profile_enable(True); c.profile_reset()
    t0 = time.time(); r = c.eigh(f, U, U / 2, beta); wall = time.time() - t0
    H0 = o.hopping_dense(o.KINDS[kind], L)
    res = orth = 0.0
    for b in range(min(B, 4)):
        H = H0 + np.diag(U * f[b] - U / 2)
        V, ev = r["evecs"][b], r["spectrum"][b]
        res = max(res, np.abs(H @ V - V * ev).max() / np.abs(ev).max())
        orth = max(orth, np.abs(V.T @ V - np.eye(n)).max())
    fam = {k: c.profile_get(k) for k in ("sytrd", "tridiag_eig", "stein", "backtransform")}
    print("%s L=%d N=%d B=%d: residual %.2e  orthogonality %.2e  wall %.3fs | " % (kind, L, n, B, res, orth, wall) +
          "  ".join("%s %.2f ms" % (k, v[0]) for k, v in fam.items()), flush=True)
    c.close()
