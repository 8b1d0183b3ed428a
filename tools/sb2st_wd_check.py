"""Column stride of the band in sb2st (one-CTA-per-SM instantiation): spectra bit for bit and the kernel's time for two builds.
python tools/sb2st_wd_check.py            (FKMC_LIB_A / FKMC_LIB_B: the two libraries; runs itself once per library)"""
import os, subprocess, sys
import numpy as np

CASES = (("cubic2d", 32, 1024), ("triangular", 31, 512), ("cubic2d", 28, 512), ("cubic2d", 26, 512), ("cubic3d", 9, 296), ("cubic2d", 24, 512))

def worker(out):
    sys.path.insert(0, ".")
    import fk_mc_b200 as fk
    res = {}
    for kind, L, B in CASES:
        c = fk.Context(kind, L, max_batch=B)
        rng = np.random.default_rng(L)
        f = (rng.random((B, c.N)) < 0.5).astype(np.int32)
        runs = [c.logz_ed(f, 3.0, 1.5, 5.0)["spectrum"].copy() for _ in range(2)]
        assert np.array_equal(runs[0], runs[1]), "run-to-run variation"
        c.profile_enable(True); c.profile_reset()
        for _ in range(3):
            c.logz_ed(f, 3.0, 1.5, 5.0)
        ms, n = c.profile_get("sb2st")
        key = "%s_%d" % (kind, L)
        res[key] = runs[0]
        res["t_" + key] = np.array([ms / max(n, 1), c.N, B])
        c.close()
    np.savez(out, **res)

if len(sys.argv) > 1:
    worker(sys.argv[1])
else:
    libs = (os.environ.get("FKMC_LIB_A", "fk_mc_b200/lib/libfkmc_b200.so"), os.environ.get("FKMC_LIB_B", "fk_mc_b200/lib_wd24/libfkmc_b200.so"))
    outs = []
    for i, lib in enumerate(libs):
        out = "/tmp/sb2st_wd_%d.npz" % i
        subprocess.check_call([sys.executable, __file__, out], env=dict(os.environ, FKMC_LIB=lib))
        outs.append(np.load(out))
    for k in [k for k in outs[0].files if not k.startswith("t_")]:
        a, b = outs[0][k], outs[1][k]
        ta, tb = outs[0]["t_" + k], outs[1]["t_" + k]
        print("%-14s N=%4d %5d matrices: identical %s, max |diff| %.1e; sb2st %.3f -> %.3f ms" % (k, int(ta[1]), a.shape[0], np.array_equal(a, b), np.abs(a - b).max(), ta[0], tb[0]), flush=True)
