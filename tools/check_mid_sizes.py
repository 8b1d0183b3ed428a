import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, fk_mc_b200 as fk, oracle_lib as o
for kind, L in [("cubic2d", 20), ("triangular", 20), ("honeycomb", 20), ("cubic2d", 28)]:
    c = fk.Context(kind, L, max_batch=300)
    N = c.N
    rng = np.random.default_rng(5)
    f = (rng.random((300, N)) < 0.5).astype(np.int32)
    r = c.logz_ed(f, 2.0, 1.0, 5.0)
    worst = 0.0
    for b in (0, 7, 299):
        ref = o.calc_ed(o.KINDS[kind], L, f[b], 2.0, 1.0, 5.0)
        worst = max(worst, np.abs(r["spectrum"][b] - ref["spectrum"]).max() / np.abs(ref["spectrum"]).max(), abs(r["logZ"][b] - ref["logZ"]) / abs(ref["logZ"]))
    print(kind, L, N, "max rel err", worst, flush=True)
    assert worst < 1e-10
