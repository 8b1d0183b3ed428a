"""Measurement-sweep observables built on the GPU eigen-decomposition.

measure_stiffness::accumulate (include/fk_mc/measures/stiffness.hpp:69-187) for hypercubic lattices: `stiffness` runs entirely on the
GPU (fkmc_stiffness_batched: eigenvectors, V^T Jm V as a DMMA GEMM, Kubo sums); `stiffness_host_contraction` keeps the numpy version of
the contractions as a cross-check (t = 1).  The reference does not register the measure at HEAD (fk_mc.hxx:101-105).
"""
import numpy as np


def _shift_index(L, ndim):
    n = L ** ndim
    idx = np.arange(n).reshape((L,) * ndim)
    left = np.roll(idx, 1, axis=0).reshape(-1)    # site with first coordinate x - 1
    right = np.roll(idx, -1, axis=0).reshape(-1)  # x + 1
    return left, right


def stiffness(ctx, f, U, mu_c, beta, ndim=2, offset=0.05, wgrid=(0.0,)):
    """Returns (stiffness [B], conductivity [B, n_w]) for configurations f [B, V]: everything on the GPU (fkmc_stiffness_batched)."""
    return ctx.stiffness(f, U, mu_c, beta, offset=offset, wgrid=wgrid)


def stiffness_host_contraction(ctx, f, U, mu_c, beta, ndim=2, offset=0.05, wgrid=(0.0,)):
    """Cross-check of fkmc_stiffness_batched: GPU eigen-decomposition, the contractions and Kubo sums in numpy."""
    r = ctx.eigh(f, U, mu_c, beta)
    B = r["spectrum"].shape[0]
    n, L = ctx.N, ctx.L
    left, right = _shift_index(L, ndim)
    wgrid = np.asarray(wgrid, dtype=np.float64)
    out, cond = np.zeros(B), np.zeros((B, len(wgrid)))
    for b in range(B):
        ev, V = r["spectrum"][b], r["evecs"][b]
        with np.errstate(over="ignore"):
            fermi = 1.0 / (1.0 + np.exp(beta * ev))
        TV = -V[left, :] - V[right, :]          # Tm V  (Tm(i, i-x) = Tm(i, i+x) = -t)
        JV = -V[left, :] + V[right, :]          # Jm V  (Jm(i, i-x) = -1, Jm(i, i+x) = +1)
        T = -np.pi * np.sum(np.einsum("ik,ik->k", V, TV) * fermi)
        mJ = V.T @ JV                           # mJ[j, i] = v_j . (Jm v_i)
        de = ev[:, None] - ev[None, :]          # e_i - e_j
        sigma = np.pi * (fermi[None, :] - fermi[:, None]) * mJ.T * mJ   # [i, j]: pi (f_j - f_i) mJ(j,i) mJ(i,j)
        mask = np.tril(np.ones((n, n), dtype=bool), -1) & (np.abs(de) > 1e-12) & (np.abs(sigma) > 1e-13)
        Vk = np.sum(2.0 * sigma[mask] / de[mask])
        out[b] = (Vk + T) / n
        for w, wv in enumerate(wgrid):
            x1 = wv - de[mask]                  # resonant terms at (e_j - e_i, sigma) and (e_i - e_j, -sigma)
            x2 = wv + de[mask]
            cond[b, w] = np.sum(offset / np.pi / (x1 * x1 + offset * offset) * sigma[mask]) \
                - np.sum(offset / np.pi / (x2 * x2 + offset * offset) * sigma[mask])
    return out, cond


def ipr(ctx, f, U, mu_c, beta):
    """measure_ipr (include/fk_mc/measures/ipr.hpp:39-56), entirely on the GPU."""
    return ctx.ipr(f, U, mu_c, beta)
