#!/bin/bash
# developer tool: 8 warps x 1 CTA/SM against 4 warps x 2 CTAs/SM of the tiled dense->band kernel at N = 256 and 512
for w in 4 8; do for L in 16 22; do
  echo "== FKMC_S1_WARPS=$w L=$L"
  FKMC_S1_WARPS=$w timeout 180 python tools/dense_time.py $L 1184 2>&1 | head -1
done; done
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, fk_mc_b200 as fk
# cubic3d L=8 -> N=512
for w in ("4", "8"):
    os.environ["FKMC_S1_WARPS"] = w
    c = fk.Context("cubic3d", 8, max_batch=1184)
    rng = np.random.default_rng(0)
    f = (rng.random((1184, c.N)) < 0.5).astype(np.int32)
    c.logz_ed(f, 4.0, 2.0, 5.0)
    c.profile_enable(True); c.profile_reset()
    for _ in range(3): c.logz_ed(f, 4.0, 2.0, 5.0)
    ms, n = c.profile_get("sy2sb")
    print("N=512 warps", w, "sy2sb %.3f ms  %.2f TFLOP/s" % (ms / n, 4 / 3 * 512 ** 3 * 1184 / (ms / n * 1e-3) * 1e-12), flush=True)
    c.close()
PY
