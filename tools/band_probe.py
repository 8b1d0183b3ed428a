"""Band path (sb2sb.cu) against the dense path: spectra, logZ and timing.  python tools/band_probe.py [B]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import fk_mc_b200 as fk

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
cases = [("cubic2d", 24), ("cubic2d", 32), ("triangular", 24), ("honeycomb", 24), ("cubic2d", 26), ("cubic2d", 16)]
if len(sys.argv) > 2:
    cases = [(sys.argv[2], int(sys.argv[3]))]
for kind, L in cases:
    c = fk.Context(kind, L, max_batch=B)
    n = c.N
    c.set_option("band_min", 16)
    rng = np.random.default_rng(5)
    fs = rng.integers(0, 2, size=(B, n)).astype(np.int32)
    U, beta = 2.0, 5.0
    res = {}
    for bp in (0, 1):
        c.set_option("band_path", bp)
        r = c.logz_ed(fs, U, U / 2, beta)
        c.profile_enable(True)
        c.profile_reset()
        t0 = time.time()
        for _ in range(3):
            r = c.logz_ed(fs, U, U / 2, beta)
        dt = (time.time() - t0) / 3
        prof = {k: c.profile_get(k)[0] / 3 for k in ("sy2sb", "sb2sb", "band_build", "sb2st", "tridiag_eig")}
        c.profile_enable(False)
        res[bp] = (r, dt, prof)
    e0, e1 = res[0][0]["spectrum"], res[1][0]["spectrum"]
    err = np.abs(e0 - e1).max()
    lz = np.abs(res[0][0]["logZ"] - res[1][0]["logZ"]).max()
    print(f"{kind} L={L} N={n} B={B}: max |eig diff| {err:.2e}  |logZ diff| {lz:.2e}  dense {res[0][1]*1e3:.1f} ms  band {res[1][1]*1e3:.1f} ms", flush=True)
    print("   dense:", {k: round(v, 2) for k, v in res[0][2].items() if v}, " band:", {k: round(v, 2) for k, v in res[1][2].items() if v}, flush=True)
    c.close()
