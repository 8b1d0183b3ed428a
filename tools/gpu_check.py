"""Quick stage-by-stage GPU diagnostics against the CPU oracle (developer tool; the formal parity
tests live in tests/).  Usage: python tools/gpu_check.py [stage ...]"""
import os
import sys
import time

import numpy as np
import scipy.linalg as sl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as o  # noqa: E402
import fk_mc_b200 as fk  # noqa: E402

stages = sys.argv[1:] or ["rng", "lattice", "tridiag", "sytrd", "ed", "kpm", "chain"]


def stage(name):
    def deco(fn):
        if name in stages:
            print("=== %s" % name, flush=True)
            try:
                fn()
            except Exception as e:  # keep going: one call should report every stage
                print("!!! %s failed: %r" % (name, e), flush=True)
        return fn
    return deco


ctx8 = fk.Context("cubic2d", 8, max_batch=64)


@stage("rng")
def _():
    for mode, V in [(0, 0), (1, 64), (1, 576), (1, 1000), (2, 0)]:
        a = ctx8.rng_stream(32167, mode, V, 3000)
        b = o.rng_stream(32167, mode, V, 3000)
        print("rng mode %d V %d: mismatches %d" % (mode, V, int((a != b).sum())))
    a = ctx8.rng_stream(-5, 0, 0, 10)
    b = o.rng_stream(-5, 0, 0, 10)
    print("rng negative seed mismatches", int((a != b).sum()))


@stage("lattice")
def _():
    for kind, L in [("cubic1d", 8), ("cubic2d", 8), ("cubic3d", 4), ("triangular", 6), ("honeycomb", 6), ("honeycomb_ref_lower", 6)]:
        c = fk.Context(kind, L)
        H = c.hopping_dense()
        Ho = o.hopping_dense(o.KINDS[kind], L)
        print("lattice %s L=%d: max|dH| %.1e sym %.1e" % (kind, L, np.abs(H - Ho).max(), np.abs(H - H.T).max()))
        c.close()


@stage("tridiag")
def _():
    rng = np.random.default_rng(1)
    for n in [5, 64, 100, 256, 576, 1024]:
        d = rng.normal(size=(3, n)) * 2
        e = rng.normal(size=(3, n - 1))
        if n >= 64:
            e[1, ::7] = 0.0
            e[2, :] *= 1e-9
        ev = ctx8.tridiag_eigvals(d, e)
        err = 0
        for b in range(3):
            ref = sl.eigvalsh_tridiagonal(d[b], e[b])
            err = max(err, np.abs(ev[b] - ref).max())
        print("tridiag n=%d: max err %.2e" % (n, err))
    # degenerate: identity and zero off-diagonals
    d = np.ones((1, 64)); e = np.zeros((1, 63))
    print("tridiag identity:", np.abs(ctx8.tridiag_eigvals(d, e) - 1).max())


@stage("sytrd")
def _():
    rng = np.random.default_rng(2)
    for n in [3, 33, 64, 100, 256, 300]:
        A = rng.normal(size=(2, n, n))
        A = A + np.transpose(A, (0, 2, 1))
        t0 = time.time()
        d, e = ctx8.sytrd(A)
        dt = time.time() - t0
        err = 0
        for b in range(2):
            ref = sl.eigvalsh(A[b])
            got = sl.eigvalsh_tridiagonal(d[b], e[b])
            err = max(err, np.abs(ref - got).max() / np.abs(ref).max())
        print("sytrd n=%d: rel err %.2e (%.3fs)" % (n, err, dt), flush=True)


@stage("ed")
def _():
    for kind, L, U, beta in [("cubic2d", 8, 1.0, 1.0), ("cubic2d", 16, 2.0, 10.0), ("cubic3d", 8, 4.0, 5.0), ("triangular", 24, 2.0, 10.0),
                             ("honeycomb", 24, 2.0, 10.0), ("honeycomb_ref_lower", 24, 2.0, 10.0), ("cubic2d", 32, 2.0, 20.0)]:
        c = fk.Context(kind, L, max_batch=4)
        n = c.N
        fs = np.stack([o.randomize_f(32167 + i, n, n // 2)[0] for i in range(3)] + [np.zeros(n, np.int32)])
        t0 = time.time()
        r = c.logz_ed(fs, U, U / 2, beta, want_caches=True)
        dt = time.time() - t0
        err = lzerr = 0
        for b in range(4):
            ref = o.calc_ed(o.KINDS[kind], L, fs[b], U, U / 2, beta)
            err = max(err, np.abs(ref["spectrum"] - r["spectrum"][b]).max() / np.abs(ref["spectrum"]).max())
            lzerr = max(lzerr, abs(ref["logZ"] - r["logZ"][b]) / max(1, abs(ref["logZ"])))
        print("ed %s L=%d N=%d: spectrum rel err %.2e logZ rel err %.2e (%.3fs)" % (kind, L, n, err, lzerr, dt), flush=True)
        c.close()


@stage("kpm")
def _():
    for kind, L, U, beta in [("cubic2d", 8, 1.0, 1.0), ("cubic2d", 16, 2.0, 10.0), ("cubic3d", 8, 4.0, 5.0), ("triangular", 24, 2.0, 10.0),
                             ("honeycomb", 24, 2.0, 10.0), ("cubic2d", 32, 2.0, 20.0)]:
        c = fk.Context(kind, L, max_batch=4)
        n = c.N
        M, G = fk.cheb_sizes(n, 2.2)
        fs = np.stack([o.randomize_f(32167 + i, n, n // 2)[0] for i in range(3)] + [np.zeros(n, np.int32)])
        t0 = time.time()
        try:
            r = c.logz_kpm(fs, U, U / 2, beta, M, G)
        except fk.FkmcError as e:
            print("kpm %s L=%d: %s" % (kind, L, e))
            c.close()
            continue
        dt = time.time() - t0
        for b in range(4):
            ref = o.calc_chebyshev(o.KINDS[kind], L, fs[b], U, U / 2, beta, M, G, emode=0)
            print("kpm %s L=%d M=%d b=%d: d_emin %.1e d_emax %.1e mom err %.1e logZ rel err %.1e" % (
                kind, L, M, b, r["e_min"][b] - ref["e_min"], r["e_max"][b] - ref["e_max"],
                np.abs(r["moments"][b] - ref["moments"]).max(), abs(r["logZ"][b] - ref["logZ"]) / abs(ref["logZ"])), flush=True)
        print("   (%.3fs)" % dt)
        c.close()


@stage("chain")
def _():
    for kind, L, U, beta, cheb, flip in [("cubic2d", 8, 1.0, 1.0, False, 0.0), ("cubic2d", 8, 4.0, 4.0, False, 0.5),
                                         ("cubic2d", 8, 4.0, 4.0, True, 0.5), ("cubic2d", 16, 2.0, 10.0, False, 0.0)]:
        nch, nsw, sl_ = 3, 4, 16
        c = fk.Context(kind, L, max_batch=nch)
        c.chain_init(nch, beta, U, mc_flip=flip, mc_reshuffle=0.1 if flip else 0.0, cheb_moves=cheb, seed=32167, sweep_len=sl_,
                     ntherm_sweeps=1, measure_energy=True, record_trace=True, max_sweeps=nsw + 1)
        c.chain_run_sweeps(nsw + 1)
        tr = c.chain_get_trace()
        se = c.chain_get_series()
        st = c.chain_get_state()
        for ch in range(nch):
            p = o.make_params(kind=o.KINDS[kind], L=L, beta=beta, U=U, mc_flip=flip, mc_reshuffle=0.1 if flip else 0.0, cheb_moves=cheb,
                              seed=32167, nsweeps=nsw, sweep_len=sl_, ntherm_sweeps=1)
            r = o.mc_run(p, rank=ch)
            t = r["trace"]
            nm = int((t["accepted"] != tr["accepted"][:, ch]).sum())
            ns = int((t["site_a"] != tr["site_a"][:, ch]).sum())
            print("chain %s L=%d cheb=%d ch=%d: accept mismatches %d site mismatches %d max|dw| %.1e max|du| %.1e dE %.1e nacc %d/%d f-diff %d" % (
                kind, L, cheb, ch, nm, ns, np.abs(t["weight"] - tr["weight"][:, ch]).max(), np.abs(t["u"] - tr["u"][:, ch]).max(),
                np.abs(r["energies"] - se["energies"][:, ch]).max(), st["naccept"][ch], r["naccept"],
                int((r["f_final"] != st["f"][ch]).sum())), flush=True)
        c.close()


def band_to_dense(AB):
    n = AB.shape[1]
    A = np.zeros((n, n))
    for d in range(AB.shape[0]):
        for c in range(n - d):
            A[c + d, c] = AB[d, c]
            A[c, c + d] = AB[d, c]
    return A


@stage("two")
def _():
    rng = np.random.default_rng(3)
    for n in [3, 9, 16, 17, 40, 64, 100, 256, 577, 1024]:
        AB = rng.normal(size=(2, 9, n))
        for d in range(9):
            AB[:, d, n - d:] = 0
        if n > 20:
            AB[1, 5:, :] = 0  # narrower band
        t0 = time.time()
        d_, e_ = ctx8.sb2st(AB)
        dt = time.time() - t0
        err = 0
        for b in range(2):
            ref = sl.eigvalsh(band_to_dense(AB[b]))
            got = sl.eigvalsh_tridiagonal(d_[b], e_[b])
            err = max(err, np.abs(ref - got).max() / np.abs(ref).max())
        print("sb2st n=%d: rel err %.2e (%.3fs)" % (n, err, dt), flush=True)
    for n in [2, 8, 9, 16, 33, 64, 100, 256, 300]:
        A = rng.normal(size=(2, n, n))
        A = A + np.transpose(A, (0, 2, 1))
        t0 = time.time()
        AB = ctx8.sy2sb(A)
        dt = time.time() - t0
        err = 0
        for b in range(2):
            ref = sl.eigvalsh(A[b])
            got = sl.eigvalsh(band_to_dense(AB[b]))
            err = max(err, np.abs(ref - got).max() / np.abs(ref).max())
        print("sy2sb n=%d: rel err %.2e (%.3fs)" % (n, err, dt), flush=True)


@stage("tiled")
def _():
    rng = np.random.default_rng(3)
    for n in [512, 520, 576, 1024]:
        A = rng.normal(size=(2, n, n))
        A = A + np.transpose(A, (0, 2, 1))
        AB = ctx8.sy2sb(A)
        err = 0
        for b in range(2):
            ref = sl.eigvalsh(A[b])
            err = max(err, np.abs(ref - sl.eigvalsh(band_to_dense(AB[b]))).max() / np.abs(ref).max())
        print("sy2sb tiled n=%d: rel err %.2e" % (n, err), flush=True)
