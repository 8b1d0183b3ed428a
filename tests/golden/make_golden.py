"""Generates tests/golden/*.json.  Run from the repo root: python tests/golden/make_golden.py

The reference (aeantipov/fk_mc) cannot be built or imported in this image, so
these vectors do NOT come from the reference binary.  They come from implementations that are
independent of both the oracle and the CUDA code:
  * spectra / logZ: numpy dense Hamiltonians built here from the lattice rules of
    src/lattice/hypercubic.cpp:116-203 + LAPACK dsyevd through scipy;
  * KPM moments / logZ: numpy dense matrix Chebyshev recursion following src/configuration.cpp:94-205
    with e_min/e_max taken from the LAPACK spectrum, and the trapezoid quadrature of
    include/fk_mc/chebyshev.hpp:36-54;
  * libstdc++ RNG streams: produced by the oracle (which calls <random> directly).
  * a short Metropolis trace from the oracle (regression pin for the RNG-consumption order).
"""
import json
import math
import os
import sys

import numpy as np
import scipy.linalg as sl

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as o  # noqa: E402


def hopping(kind, L, t=1.0, tp=1.0):
    """Independent numpy statement of the lattice rules (last coordinate fastest)."""
    nd = {"cubic1d": 1, "cubic2d": 2, "cubic3d": 3, "triangular": 2, "honeycomb": 2}[kind]
    N = L ** nd
    H = np.zeros((N, N))
    for i in range(N):
        pos = list(np.unravel_index(i, (L,) * nd))
        if kind.startswith("cubic") or kind == "triangular":
            for d in range(nd):
                for s in (-1, 1):
                    q = list(pos)
                    q[d] = (q[d] + s) % L
                    H[i, np.ravel_multi_index(q, (L,) * nd)] += -t
            if kind == "triangular":
                for s in (-1, 1):
                    q = [(pos[0] + s) % L, (pos[1] + s) % L]
                    H[i, np.ravel_multi_index(q, (L, L))] += -tp
        else:  # brick-wall honeycomb: x = last coordinate, A <=> (x+y) even hops up (y+1)
            y, x = pos
            H[i, np.ravel_multi_index([y, (x - 1) % L], (L, L))] += -t
            H[i, np.ravel_multi_index([y, (x + 1) % L], (L, L))] += -t
            yy = (y + 1) % L if (x + y) % 2 == 0 else (y - 1) % L
            H[i, np.ravel_multi_index([yy, x], (L, L))] += -t
    return H


def kpm_numpy(H, beta, M, G):
    N = H.shape[0]
    ev = sl.eigh(H, eigvals_only=True, driver="evd")
    e_min, e_max = ev[0], ev[-1]
    a, b = (e_max - e_min) / 2, (e_max + e_min) / 2
    X = (H - b * np.eye(N)) / a
    T0, T1 = np.eye(N), X.copy()
    mom = np.zeros(M)
    is_set = [False] * M
    mom[0], mom[1] = 1.0, np.trace(X) / N
    is_set[0] = is_set[1] = True
    for m in range(2, M // 2 + 1):
        T0, T1 = T1, 2 * X @ T1 - T0
        if not is_set[m]:
            mom[m] = np.trace(T1) / N
            is_set[m] = True
        k = 2 * m - 1
        if k < M and k >= M // 2:
            mom[k] = (2 * np.trace(T0 @ T1) - np.trace(X)) / N
            is_set[k] = True
            if k != M - 1:
                mom[k + 1] = 2 * np.trace(T1 @ T1) / N - 1
                is_set[k + 1] = True
    theta = np.linspace(0.0, 1.0, G)
    x = -np.cos(np.pi * theta)
    F = N * np.log(1 + np.exp(-beta * (a * x + b)))

    def moment(order):
        T = np.cos(order * np.arccos(x))
        return 0.5 * np.sum((F[1:] * T[1:] + F[:-1] * T[:-1]) * np.diff(theta))

    logz = moment(0) + sum(2 * moment(m) * mom[m] for m in range(1, M))
    return dict(e_min=e_min, e_max=e_max, moments=mom.tolist(), logZ=logz)


def main():
    cases = []
    for kind, L, U, beta, seed in [("cubic2d", 8, 1.0, 1.0, 32167), ("cubic2d", 16, 2.0, 10.0, 32167), ("cubic3d", 4, 4.0, 5.0, 7),
                                   ("cubic3d", 8, 4.0, 5.0, 32167), ("triangular", 6, 2.0, 10.0, 11), ("honeycomb", 6, 2.0, 10.0, 13),
                                   ("triangular", 24, 2.0, 10.0, 32167), ("honeycomb", 24, 2.0, 10.0, 32167), ("cubic1d", 12, 1.5, 3.0, 5)]:
        H0 = hopping(kind, L)
        N = H0.shape[0]
        f, _ = o.randomize_f(seed, N, N // 2)
        H = H0 + np.diag(U * f - U / 2)
        ev = sl.eigh(H, eigvals_only=True, driver="evd")
        logz = float(np.sum(np.log1p(np.exp(-beta * ev))))
        M = int(math.log(N) * 2.2)
        M += M % 2
        G = max(2 * M, 10)
        case = dict(kind=kind, L=L, U=U, mu_c=U / 2, beta=beta, seed=seed, f=f.tolist(), spectrum=ev.tolist(), logZ=logz, M=M, G=G)
        if N <= 600:
            case["kpm"] = kpm_numpy(H, beta, M, G)
        cases.append(case)
    json.dump(dict(cases=cases), open(os.path.join(HERE, "spectra.json"), "w"))

    rng = dict(seed=32167,
               raw=[int(x) for x in o.rng_stream(32167, 0, 0, 40)],
               uniform_int_64=[int(x) for x in o.rng_stream(32167, 1, 64, 40)],
               uniform_int_576=[int(x) for x in o.rng_stream(32167, 1, 576, 40)],
               uniform_real=[float(x) for x in o.rng_stream(32167, 2, 0, 40)],
               randomize_f_8x8=o.randomize_f(32167, 64, 32)[0].tolist())
    json.dump(rng, open(os.path.join(HERE, "rng.json"), "w"))

    traces = []
    for cheb, flip, resh, U, beta in [(False, 0.0, 0.0, 1.0, 1.0), (False, 0.5, 0.1, 4.0, 4.0), (True, 0.5, 0.1, 4.0, 4.0)]:
        p = o.make_params(kind=o.CUBIC2D, L=8, beta=beta, U=U, mc_flip=flip, mc_reshuffle=resh, cheb_moves=cheb, seed=32167, nsweeps=3,
                          sweep_len=16, ntherm_sweeps=1)
        r = o.mc_run(p, rank=1)
        t = r["trace"]
        traces.append(dict(cheb=cheb, mc_flip=flip, mc_reshuffle=resh, U=U, beta=beta, rank=1, move=t["move"].tolist(),
                           site_a=t["site_a"].tolist(), site_b=t["site_b"].tolist(), accepted=t["accepted"].tolist(),
                           weight=t["weight"].tolist(), u=t["u"].tolist(), energies=r["energies"].tolist(),
                           d2energies=r["d2energies"].tolist(), f_final=r["f_final"].tolist()))
    json.dump(dict(traces=traces), open(os.path.join(HERE, "mc_trace.json"), "w"))
    print("wrote spectra.json, rng.json, mc_trace.json")


if __name__ == "__main__":
    main()
