"""sy2sb time vs number of busy SMs (shared-resource vs per-SM bound): python tools/perf_sy2sb_occ.py [L]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fk_mc_b200 as fk
L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
c = fk.Context("cubic2d", L, max_batch=296)
rng = np.random.default_rng(0)
f = (rng.random((296, c.N)) < 0.5).astype(np.int32)
c.logz_ed(f[:148], 2.0, 1.0, 10.0)
c.profile_enable(True)
for B in (1, 18, 37, 74, 111, 148, 296):
    c.profile_reset()
    for _ in range(2):
        c.logz_ed(f[:B], 2.0, 1.0, 10.0)
    ms, n = c.profile_get("sy2sb")
    print("B=%3d sy2sb %.3f ms/launch -> %.2f TFLOP/s" % (B, ms / n, 4 / 3 * c.N ** 3 * B / (ms / n * 1e-3) * 1e-12), flush=True)
