"""Kernel-family timings on the GPU (developer tool).  Usage: python tools/gpu_perf.py [case ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fk_mc_b200 as fk  # noqa: E402

cases = sys.argv[1:] or ["ed64", "ed256", "ed512", "ed1024", "kpm1024", "kpm256"]
CFG = {
    "ed64": ("cubic2d", 8, 4096, 1.0, 1.0, False),
    "ed256": ("cubic2d", 16, 4096, 2.0, 10.0, False),
    "ed512": ("cubic3d", 8, 1024, 4.0, 5.0, False),
    "ed576": ("triangular", 24, 1024, 2.0, 10.0, False),
    "ed1024": ("cubic2d", 32, 592, 2.0, 20.0, False),
    "kpm1024": ("cubic2d", 32, 1184, 2.0, 20.0, True),
    "kpm256": ("cubic2d", 16, 4096, 2.0, 10.0, True),
    "kpm576": ("triangular", 24, 1184, 2.0, 10.0, True),
}
for name in cases:
    kind, L, B, U, beta, cheb = CFG[name]
    c = fk.Context(kind, L, max_batch=B)
    if os.environ.get("FKMC_TRIDIAG"):
        c.set_option("tridiag", int(os.environ["FKMC_TRIDIAG"]))
    N = c.N
    sweep_len = 4
    c.chain_init(B, beta, U, cheb_moves=cheb, sweep_len=sweep_len, ntherm_sweeps=1000, measure_energy=False, max_sweeps=8)
    c.chain_run_sweeps(1)
    c.sync()
    c.profile_enable(True)
    c.profile_reset()
    t0 = time.time()
    c.chain_run_sweeps(2)
    c.sync()
    wall = time.time() - t0
    props = 2 * sweep_len * B
    line = "%s N=%d B=%d: %.1f proposals/s (wall %.3fs)" % (name, N, B, props / wall, wall)
    for fam in ["build_h", "sytrd", "sy2sb", "sb2st", "tridiag_eig", "kpm", "chain_step"]:
        ms, n = c.profile_get(fam)
        if n:
            line += " | %s %.3f ms/launch" % (fam, ms / n)
    print(line, flush=True)
    if not cheb:
        ms, n = c.profile_get("sytrd")
        if n == 0:
            ms1, n = c.profile_get("sy2sb")
            ms2, n2 = c.profile_get("sb2st")
            print("   sy2sb: %.2f TFLOP/s (4/3 N^3), %.0f matrices/s; sb2st %.0f matrices/s" % (
                4.0 / 3.0 * N ** 3 * B / (ms1 / n * 1e-3) * 1e-12, B / (ms1 / n * 1e-3), B / (ms2 / n2 * 1e-3)))
            ms = ms1 + ms2
        fl = 4.0 / 3.0 * N ** 3 * B / (ms / n * 1e-3)
        print("   tridiagonalisation: %.2f TFLOP/s (4/3 N^3), %.0f matrices/s" % (fl * 1e-12, B / (ms / n * 1e-3)))
        ms, n = c.profile_get("tridiag_eig")
        print("   tridiag_eig: %.0f matrices/s" % (B / (ms / n * 1e-3)))
    else:
        ms, n = c.profile_get("kpm")
        M, G = fk.cheb_sizes(N)
        bk = 24.0 * N * N * (M / 2 - 1) + 16.0 * N * N
        print("   kpm: %.0f proposals/s, algorithmic %.1f GB/s" % (B / (ms / n * 1e-3), bk * B / (ms / n * 1e-3) * 1e-9))
    c.close()
