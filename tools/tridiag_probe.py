"""tridiag_eig_kernel: time per launch and accuracy against LAPACK on the c5 / c2 shapes.  FKMC_LIB selects a developer build.
usage: python tools/tridiag_probe.py"""
import sys
import numpy as np
sys.path.insert(0, ".")
import fk_mc_b200 as fk

for kind, L, B, U, beta in (("cubic2d", 32, 1024, 2.0, 20.0), ("cubic2d", 16, 4096, 2.0, 10.0), ("cubic3d", 8, 1024, 4.0, 5.0)):
    c = fk.Context(kind, L, max_batch=B)
    n = c.N
    rng = np.random.default_rng(11)
    fs = rng.integers(0, 2, size=(B, n)).astype(np.int32)
    fs[0] = 0
    fs[1] = 1
    fs[2] = (np.arange(n) // L + np.arange(n)) % 2  # checkerboard: degenerate spectrum
    r = c.logz_ed(fs, U, U / 2, beta)
    c.profile_enable(True)
    c.profile_reset()
    for _ in range(5):
        r = c.logz_ed(fs, U, U / 2, beta)
    ms, cnt = c.profile_get("tridiag_eig")
    c.profile_enable(False)
    H0 = c.hopping_dense()
    err = 0.0
    for b in list(range(8)) + [B - 1]:
        H = H0 + np.diag(U * fs[b] - U / 2)
        w = np.linalg.eigvalsh(H)
        err = max(err, np.abs(w - r["spectrum"][b]).max() / max(1.0, np.abs(w).max()))
    print(f"{kind} L={L} N={n} B={B}: tridiag_eig {ms / cnt:.3f} ms per launch, max rel eigenvalue error vs LAPACK {err:.2e}, "
          f"checksum {r['spectrum'].sum():.12e} logZ sum {r['logZ'].sum():.12e}", flush=True)
    # worst case for the grouped recurrences: every matrix with a massively degenerate spectrum (ordered state), whose tridiagonal
    # has runs of negligible off-diagonal entries
    fo = np.tile(fs[2], (B, 1))
    fo[:, 0] = rng.integers(0, 2, size=B)
    r = c.logz_ed(fo, U, U / 2, beta)
    c.profile_enable(True)
    c.profile_reset()
    for _ in range(3):
        r = c.logz_ed(fo, U, U / 2, beta)
    ms, cnt = c.profile_get("tridiag_eig")
    c.profile_enable(False)
    w = np.linalg.eigvalsh(H0 + np.diag(U * fo[5] - U / 2))
    print(f"    ordered (checkerboard) batch: tridiag_eig {ms / cnt:.3f} ms per launch, error {np.abs(w - r['spectrum'][5]).max():.2e}", flush=True)
    c.close()
