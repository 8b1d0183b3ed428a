"""Bit-for-bit comparison of the band -> tridiagonal stage at sweep lag 2 (default build) and lag 3 (make VARIANT=lag3 EXTRA=-DFKMC_SB2ST_LAG=3):
a data race of the tighter pipeline would show as different or run-to-run varying spectra.
python tools/sb2st_lag_check.py            (runs itself once per library in a subprocess)"""
import os, subprocess, sys
import numpy as np

def worker(out):
    sys.path.insert(0, ".")
    import fk_mc_b200 as fk
    res = {}
    for kind, L, B in (("cubic2d", 16, 2048), ("cubic2d", 24, 1024), ("triangular", 24, 512), ("cubic2d", 32, 1024), ("cubic3d", 8, 512), ("cubic2d", 20, 512)):
        c = fk.Context(kind, L, max_batch=B)
        rng = np.random.default_rng(L)
        f = (rng.random((B, c.N)) < 0.5).astype(np.int32)
        runs = [c.logz_ed(f, 3.0, 1.5, 5.0)["spectrum"].copy() for _ in range(3)]
        assert all(np.array_equal(runs[0], r) for r in runs[1:]), "run-to-run variation"
        res["%s_%d" % (kind, L)] = runs[0]
        c.close()
    np.savez(out, **res)

if len(sys.argv) > 1:
    worker(sys.argv[1])
else:
    outs = []
    other = os.environ.get("FKMC_LIB_B", "fk_mc_b200/lib_lag3/libfkmc_b200.so")  # any other developer build to compare bit for bit
    for tag, lib in (("lag2", "fk_mc_b200/lib/libfkmc_b200.so"), ("lag3", other)):
        out = "/tmp/sb2st_%s.npz" % tag
        subprocess.check_call([sys.executable, __file__, out], env=dict(os.environ, FKMC_LIB=lib))
        outs.append(np.load(out))
    for k in outs[0].files:
        a, b = outs[0][k], outs[1][k]
        print("%-14s %5d matrices: identical %s, max |diff| %.1e" % (k, a.shape[0], np.array_equal(a, b), np.abs(a - b).max()))
