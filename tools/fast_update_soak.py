"""Long fast-update runs at the BASELINE sizes: tracked spectra against a fresh eigensolve after many sweeps (refresh checks every 64 sweeps
inside the run raise FKMC_ERR_NOCONV on a deviation > 1e-10).  python tools/fast_update_soak.py [sweeps]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fk_mc_b200 as fk
nsw = int(sys.argv[1]) if len(sys.argv) > 1 else 150
for kind, L, n, beta, U, flip in [("cubic2d", 16, 512, 10.0, 2.0, 0.3), ("cubic3d", 8, 256, 5.0, 4.0, 0.0), ("triangular", 24, 256, 10.0, 2.0, 0.3),
                                  ("honeycomb", 24, 128, 10.0, 2.0, 0.0), ("cubic2d", 32, 64, 20.0, 2.0, 0.2)]:
    c = fk.Context(kind, L, max_batch=n)
    c.chain_init(n, beta, U, mc_flip=flip, seed=777, sweep_len=16, ntherm_sweeps=0, max_sweeps=nsw, fast_update=True, fu_refresh_sweeps=nsw + 1000)
    t0 = time.time()
    c.chain_run_sweeps(nsw)          # no refresh at all: nsw * 16 proposals of pure rank-one tracking
    dt = time.time() - t0
    st = c.chain_get_state(spectrum=True)
    fresh = c.logz_ed(st["f"], U, U / 2, beta)
    dev = np.abs(st["spectrum"] - fresh["spectrum"]).max() / np.abs(fresh["spectrum"]).max()
    dlz = np.abs(st["logZ"] - fresh["logZ"]).max() / np.abs(fresh["logZ"]).max()
    acc = st["naccept"].sum() / float(n * 16 * nsw)
    print("%-10s L=%2d N=%4d chains=%3d: %d sweeps without refresh, %5.1f accepted moves per chain, max |d eps| / max|eps| = %.2e, max |d logZ| / |logZ| = %.2e, %.0f proposals/s"
          % (kind, L, c.N, n, nsw, st["naccept"].mean(), dev, dlz, n * 16 * nsw / dt), flush=True)
    assert dev < 1e-10 and dlz < 1e-10, "tracked spectrum drifted"
    c.close()
