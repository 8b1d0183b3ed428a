"""sb2st time against the warp count per CTA (option sb2st_warps; 0 = the launcher's choice).  python tools/sb2st_warps_scan.py"""
import sys
sys.path.insert(0, '.')
import numpy as np, fk_mc_b200 as fk
for L, B, ws in ((16, 4096, (0, 2, 3, 4, 5, 6, 8)), (24, 1024, (0, 4, 6, 8, 9, 10, 12)), (32, 1024, (0, 8, 10, 12, 14, 16))):
    c = fk.Context("cubic2d", L, max_batch=B)
    rng = np.random.default_rng(0)
    f = (rng.random((B, c.N)) < 0.5).astype(np.int32)
    for w in ws:
        try:
            c.set_option("sb2st_warps", w)
        except Exception as e:
            print("L=%d warps=%d rejected" % (L, w)); continue
        c.logz_ed(f, 2.0, 1.0, 10.0)
        c.profile_enable(True); c.profile_reset()
        for _ in range(3):
            c.logz_ed(f, 2.0, 1.0, 10.0)
        ms, n = c.profile_get("sb2st")
        c.profile_enable(False)
        print("L=%d N=%d warps=%d sb2st %.3f ms per %d matrices" % (L, c.N, w, ms / n, B), flush=True)
    c.close()
