"""Developer timing of the dense stages: python tools/dense_time.py [L] [B] -> ms per launch of sy2sb / sb2st / tridiag_eig."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fk_mc_b200 as fk
L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
c = fk.Context("cubic2d", L, max_batch=B)
rng = np.random.default_rng(0)
f = (rng.random((B, c.N)) < 0.5).astype(np.int32)
c.logz_ed(f, 2.0, 1.0, 10.0)
c.profile_enable(True); c.profile_reset()
for _ in range(3):
    r = c.logz_ed(f, 2.0, 1.0, 10.0)
out = []
for k in ("build_h", "band_build", "sy2sb", "sb2sb", "sb2st", "tridiag_eig"):
    ms, n = c.profile_get(k)
    if n:
        out.append("%s %.3f ms" % (k, ms / n))
N = c.N
ms, n = c.profile_get("sy2sb")  # (not launched when the band path applies: set_option("band_path", 0) for the dense reduction)
tail = "  | sy2sb %.2f TFLOP/s" % (4 / 3 * N ** 3 * B / (ms / n * 1e-3) * 1e-12) if n else ""
print("L=%d B=%d: " % (L, B) + "  ".join(out) + tail, flush=True)
print("logZ[0:2] =", r["logZ"][:2])
