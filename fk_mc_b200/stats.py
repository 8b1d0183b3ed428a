"""Host-side error analysis of the Monte Carlo series: log-binning, jackknife, plateau-bin estimate, specific heat.

Mirrors include/fk_mc/binning.hpp:89-171 (calc_stats, bin<D>, accumulate_binning, calc_cor_length),
include/fk_mc/jackknife.hpp:50-82 (jack, accumulate_jackknife), prog/data_save.hpp:108-122 (estimate_bin) and
prog/data_save.hxx:158-199 (cv = beta^2 (<E^2> - <d2E> - <E>^2) / N).  A bin-stats row is (n, mean, variance, stderr)
exactly like the reference's bin_stats_t; /stats/<obs> of the reference's HDF5 layout is that 4-vector.
"""
import math

import numpy as np

MAX_BIN_DEPTH = 15  # BINNING_RANGE in include/fk_mc/binning.hpp:18


def pool_chains(series):
    """Series of several chains -> one 1-D series, CHAIN-MAJOR: chain 0's measurements in Monte Carlo time order, then chain 1's, ...

    The library returns per-chain series as [measurement][chain] (fkmc_chain_get_series).  The reference gathers the ranks one
    after the other (src/measures/energy.cpp:32-47), so its bins run along MC time inside one rank; flattening [measurement][chain]
    row by row would instead average neighbouring *independent* chains in the shallow bin levels and hide the autocorrelation.
    A 1-D input is one chain (or an already pooled series) and is returned as is."""
    a = np.asarray(series, dtype=np.float64)
    if a.ndim == 1:
        return a
    if a.ndim != 2:
        raise ValueError("series must be 1-D (one chain / already pooled) or 2-D [measurement][chain]")
    return np.ascontiguousarray(a.T).reshape(-1)


def calc_stats(x):
    """(n, mean, unbiased variance, sqrt(variance / n)) -- binning.hpp:89-96."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    mean = x.sum() / n
    var = ((x - mean) ** 2).sum() / (n - 1) if n > 1 else float("nan")
    return (n, float(mean), float(var), float(math.sqrt(var / n)) if n > 1 else float("nan"))


def bin_series(x, depth):
    """Averages of 2^depth consecutive samples, incomplete tail dropped (binned_iterator, binning.hpp:26-44)."""
    x = np.asarray(x, dtype=np.float64)
    step = 1 << depth
    if step > x.size:
        raise ValueError("Can't bin with binning step(%d)> container size (%d)" % (step, x.size))  # binning.hpp:66-68
    n = x.size // step
    return x[: n * step].reshape(n, step).mean(axis=1)


def bin_stats(x, depth):
    return calc_stats(bin_series(x, depth))


def accumulate_binning(x, max_depth):
    """Rows for depth 0..max_depth (binning_accumulator, binning.hpp:100-128)."""
    if max_depth > MAX_BIN_DEPTH:
        raise ValueError("bin_depth =%d> compiled bin size" % max_depth)
    return [bin_stats(x, d) for d in range(max_depth + 1)]


def calc_cor_length(rows):
    """tau_i = (2^i var_i / var_0 - 1) / 2 -- binning.hpp:163-171."""
    s0 = rows[0][2]
    return [0.5 * ((2.0 ** i) * r[2] / s0 - 1.0) for i, r in enumerate(rows)]


def jack(F, series, depth=0):
    """Jackknife of F(<x_1>, <x_2>, ...) over binned series -- jackknife.hpp:50-82.  Returns (n, value, variance, stderr)."""
    data = [bin_series(s, depth) for s in series]
    n = data[0].size
    means = np.array([d.sum() / n for d in data])
    u0 = F(*means)
    u = np.empty(n)
    for j in range(n):
        loo = [(n * means[i] - data[i][j]) / (n - 1) for i in range(len(data))]
        u[j] = F(*loo)
    _, ubar, _, uerr = calc_stats(u)
    u_avg = u0 - (n - 1) * (ubar - u0)
    du = (n - 1) * uerr
    return (n, float(u_avg), float(du * du * n), float(du))


def accumulate_jackknife(F, series, max_depth):
    return [jack(F, series, d) for d in range(max_depth + 1)]


def estimate_bin(rows):
    """Index of the bin level where the error bar has saturated -- prog/data_save.hpp:108-122."""
    errors = np.array([r[3] for r in rows], dtype=np.float64)
    rel_error, f, ind = 1.0, True, len(errors) - 1
    while f and ind > 0:
        with np.errstate(divide="ignore", invalid="ignore"):  # C++ semantics: x/0 is inf or nan, the comparisons then fail
            cur = abs(errors[ind - 1] / errors[ind] - 1.0)
        f = cur < 0.05 and cur < rel_error
        rel_error = cur if f else rel_error
        if f:
            ind -= 1
    return ind


def max_bin_depth(n_samples, min_bins=4):
    """Deepest level that still leaves `min_bins` bins, capped at the reference's compile-time depth."""
    d = 0
    while d < MAX_BIN_DEPTH and (n_samples >> (d + 1)) >= min_bins:
        d += 1
    return d


def energy_report(energies, d2energies, beta, volume, max_depth=None):
    """save_energy (prog/data_save.hxx:158-199): binning of E and d2E, jackknife of the specific heat.  The reference bins the
    series in reverse order (rbegin..rend); so do we.  2-D input is [measurement][chain] and is pooled chain-major (pool_chains)."""
    e = pool_chains(energies)[::-1]
    d2 = pool_chains(d2energies)[::-1]
    if max_depth is None:
        max_depth = max_bin_depth(e.size)
    out = {}
    for name, x in (("energy", e), ("d2energy", d2)):
        rows = accumulate_binning(x, max_depth)
        b = estimate_bin(rows)
        out[name] = dict(binning=rows, cor_length=calc_cor_length(rows), bin=b, stats=rows[b])
    cv_rows = accumulate_jackknife(lambda a, a2, de2: beta * beta * (a2 - de2 - a * a) / volume, [e, e * e, d2], max_depth)
    b = estimate_bin(cv_rows)
    out["cv"] = dict(binning=cv_rows, cor_length=calc_cor_length(cv_rows), bin=b, stats=cv_rows[b])
    return out


def fstats_report(nf0, nfpi, max_depth=None):
    """save_fstats (prog/data_save.hxx:200-236): binning of n_f(q=0), n_f(q=pi), jackknife of the f-susceptibilities <n^2> - <n>^2 and of the
    Binder cumulants 1 - <n^4> / (3 <n^2>^2) at the bin level picked for fsusc_0.  Input: 1-D series or [measurement][chain] (pooled chain-major);
    reversed like the reference (rbegin..rend) for the binning rows."""
    n0, npi = pool_chains(nf0), pool_chains(nfpi)
    if max_depth is None:
        max_depth = max_bin_depth(n0.size)
    out = {}
    for name, x in (("nf_0", n0), ("nf_pi", npi)):
        rows = accumulate_binning(x[::-1], max_depth)
        b = estimate_bin(rows)
        out[name] = dict(binning=rows, cor_length=calc_cor_length(rows), bin=b, stats=rows[b])
    disp = lambda x, x2: x2 - x * x  # noqa: E731
    for name, x in (("fsusc_0", n0), ("fsusc_pi", npi)):
        rows = accumulate_jackknife(disp, [x, x * x], max_depth)   # the reference passes these series un-reversed
        b = estimate_bin(rows)
        out[name] = dict(binning=rows, cor_length=calc_cor_length(rows), bin=b, stats=rows[b])
    nf_bin = out["fsusc_0"]["bin"]
    binder = lambda x2, x4: 1.0 - x4 / 3.0 / x2 / x2  # noqa: E731
    out["binder_0"] = dict(bin=nf_bin, stats=jack(binder, [n0 ** 2, n0 ** 4], nf_bin))
    out["binder_pi"] = dict(bin=nf_bin, stats=jack(binder, [npi ** 2, npi ** 4], nf_bin))
    return out


def _history_rows(hist):
    """Per-measurement rows [n_meas_total, N] from a history: [measurement][N] (one chain) or [measurement][chain][N] (pooled chain-major,
    like the reference's gathered ranks)."""
    a = np.asarray(hist, dtype=np.float64)
    if a.ndim == 3:
        a = np.transpose(a, (1, 0, 2)).reshape(-1, a.shape[2])
    if a.ndim != 2:
        raise ValueError("history must be [measurement][N] or [measurement][chain][N]")
    return a


def dos_report(spectrum_history, wgrid, offset, beta, max_depth=None):
    """save_glocal (prog/data_save.hxx:265-345): local density of states dos(w) = -Im sum_k 1 / (w - eps_k + i offset) / (pi N) of every measured
    spectrum; binning of dos(0) -> `dos0`; `dos_err` = rows [w, mean, stderr] at the bin level picked for dos0; nc = integral of dos(w) f(w)
    over the grid (trapezoid) with its error.  spectrum_history: [measurement][N] or [measurement][chain][N]."""
    sp = _history_rows(spectrum_history)
    n_meas, vol = sp.shape
    wgrid = np.asarray(wgrid, dtype=np.float64)

    def dos_at(w):
        return (offset / ((w - sp) ** 2 + offset * offset)).sum(axis=1) / (math.pi * vol)   # == -Im(sum 1/(w - eps + i offset)) / (pi N)

    if max_depth is None:
        max_depth = max_bin_depth(n_meas)
    rows0 = accumulate_binning(dos_at(0.0)[::-1], max_depth)
    b = estimate_bin(rows0)
    table = np.zeros((len(wgrid), 3))
    for i, w in enumerate(wgrid):
        st = bin_stats(dos_at(w)[::-1], b)
        table[i] = (w, st[1], st[3])
    out = dict(dos0=dict(binning=rows0, cor_length=calc_cor_length(rows0), bin=b, stats=rows0[b]), dos_err=table)
    if len(wgrid) > 1:
        with np.errstate(over="ignore"):
            fermi = 1.0 / (1.0 + np.exp(beta * wgrid))
        trapz = getattr(np, "trapezoid", None) or np.trapz
        nc = float(trapz(table[:, 1] * fermi, wgrid))
        nc_err = float(math.sqrt(trapz((table[:, 2] * fermi) ** 2, wgrid)))
        out["nc"] = (float("nan"), nc, float("nan"), nc_err)   # save_bin_data row: only mean and error are set by the reference
    return out


def ipr_report(spectrum_history, ipr_history, wgrid, offset, max_depth=None):
    """save_ipr (prog/data_save.hxx:487-532): Lorentzian-weighted inverse participation ratio
    ipr(w) = sum_k L(w - eps_k) ipr_k^4 / sum_k L(w - eps_k)  (the measure stores the L4 NORM, hence the 4th power); binning of ipr(0) -> `ipr0`;
    `ipr_err` = rows [w, mean, stderr] at the bin level picked for ipr0."""
    sp, ip = _history_rows(spectrum_history), _history_rows(ipr_history)
    if sp.shape != ip.shape:
        raise ValueError("spectrum_history and ipr_history differ in shape")
    wgrid = np.asarray(wgrid, dtype=np.float64)
    ip4 = ip ** 4

    def ipr_at(w):
        lor = offset / ((w - sp) ** 2 + offset * offset)
        return (lor * ip4).sum(axis=1) / lor.sum(axis=1)

    if max_depth is None:
        max_depth = max_bin_depth(sp.shape[0])
    rows0 = accumulate_binning(ipr_at(0.0)[::-1], max_depth)
    b = estimate_bin(rows0)
    table = np.zeros((len(wgrid), 3))
    for i, w in enumerate(wgrid):
        st = bin_stats(ipr_at(w)[::-1], b)
        table[i] = (w, st[1], st[3])
    return dict(ipr0=dict(binning=rows0, cor_length=calc_cor_length(rows0), bin=b, stats=rows0[b]), ipr_err=table)


def fcorrel_report(focc_history, dims, max_depth=None):
    """save_fcorrel (prog/data_save.hxx:347-420): f-electron density correlations along the lattice axes,
        C(l) = jackknife over bins of  sum_i sum_d dev_i (dev_{i - l e_d} + dev_{i + l e_d}) / (2 D V),   dev_i = <f_i>_bin - nf_mean_i
    with nf_mean_i the binned mean of site i at the bin level picked from site 0's series; the bin level for all l is then the one picked for
    C(0).  Rows of `fcorrel`: [l, C(l), error, C(l) / C(0), error of the ratio], l = 0 .. L/2; `fcorrel_q` = forward DFT of the symmetrised C(r).
    focc_history: [measurement][V] or [measurement][chain][V]."""
    fo = _history_rows(focc_history)            # [n_total][V]
    n_tot, vol = fo.shape
    dims = tuple(int(d) for d in dims)
    if int(np.prod(dims)) != vol:
        raise ValueError("fcorrel_report: dims do not match the history")
    if max_depth is None:
        max_depth = max_bin_depth(n_tot)
    series = [fo[::-1, i] for i in range(vol)]  # reversed like the reference (rbegin..rend)
    nf_bin = estimate_bin(accumulate_binning(series[0], max_depth))
    nf_mean = np.array([bin_stats(x, nf_bin)[1] for x in series])

    def fcorrel_f(l):
        def F(*means):
            dev = (np.asarray(means) - nf_mean).reshape(dims)
            out = 0.0
            for d in range(len(dims)):
                out += (dev * (np.roll(dev, l, axis=d) + np.roll(dev, -l, axis=d))).sum()
            return out / vol / (2.0 * len(dims))
        return F

    c0_rows = accumulate_jackknife(fcorrel_f(0), series, max_depth)
    nf_bin = estimate_bin(c0_rows)
    c0_mean, c0_err = c0_rows[nf_bin][1], c0_rows[nf_bin][3]
    half = dims[0] // 2
    table = np.zeros((half + 1, 5))
    cr = np.zeros(dims[0])
    per_l = {}
    for l in range(half + 1):
        st = jack(fcorrel_f(l), series, nf_bin)
        per_l["fcorrel_%d" % l] = st
        with np.errstate(divide="ignore", invalid="ignore"):
            table[l] = (l, st[1], st[3], st[1] / c0_mean,
                        math.sqrt((st[3] / c0_mean) ** 2 + (st[1] / (c0_mean * c0_mean) * c0_err) ** 2) if c0_mean != 0 else float("nan"))
        cr[l] = st[1]
        if l > 0:
            cr[dims[0] - l] = st[1]
    return dict(fcorrel=table, fcorrel_q=np.fft.fft(cr), bin=nf_bin, per_l=per_l)


def _wstring(w):
    """The reference's dataset suffix: std::to_string(float(Re w)) and (Im w) with trailing zeros erased, joined by '_' (data_save.hxx:566-571)."""
    def one(x):
        return ("%f" % np.float32(x)).rstrip("0")
    return one(w.real) + "_" + one(w.imag)


def gwr_report(eigenfunctions_history, spectrum_history, wgrid, imag_offset, dims, save_only_dos=False):
    """save_gwr (prog/data_save.hxx:535-709), 2-D lattices: the measurement-averaged Green's function
        G(w; r1, r2) = < V (w + i xi - eps)^-1 V^T >
    from the eigenfunction and spectrum histories; per frequency the average and the typical (geometric-mean) local density of states
    (`tdos` row: Re w, Im w, dos_geom, 0, dos, 0, dos_geom / dos, 0), the full matrix `gr_full`, its translation average G(w; r1 - r2)
    `gr` [L0][L1] and the lattice Fourier transform `gk` (forward DFT, as FFTW_FORWARD).
    eigenfunctions_history: [measurement][N][N] or [measurement][chain][N][N] with [..., i, k] = component i of eigenvector k;
    spectrum_history: [measurement][N] or [measurement][chain][N].  Returns dict(tdos_gwr [n_w, 8], per_w {wstring: {...}})."""
    ev = np.asarray(eigenfunctions_history, dtype=np.float64)
    if ev.ndim == 4:
        ev = np.transpose(ev, (1, 0, 2, 3)).reshape(-1, ev.shape[2], ev.shape[3])
    sp = _history_rows(spectrum_history)
    n_meas, vol = sp.shape
    if ev.shape != (n_meas, vol, vol) or int(np.prod(dims)) != vol or len(dims) != 2:
        raise ValueError("gwr_report: histories / dims mismatch (2-D lattices only)")
    L0, L1 = int(dims[0]), int(dims[1])
    idx = np.arange(vol)
    p0, p1 = idx // L1, idx % L1                                    # index_to_pos: last coordinate fastest
    d0 = (L0 + p0[None, :] - p0[:, None]) % L0                      # r_j - r_i
    d1 = (L1 + p1[None, :] - p1[:, None]) % L1
    tdos = np.zeros((len(wgrid), 8))
    per_w = {}
    for wi, w0 in enumerate(wgrid):
        w = complex(w0) + 1j * imag_offset
        g_re, g_im = np.zeros((vol, vol)), np.zeros((vol, vol))
        for m in range(n_meas):
            wme = w.real - sp[m]
            den = 1.0 / (wme * wme + w.imag * w.imag)
            g_im += (ev[m] * (-w.imag * den)) @ ev[m].T / n_meas
            g_re += (ev[m] * (wme * den)) @ ev[m].T / n_meas
        ldos = np.diag(g_im) / (-math.pi)
        dos_val = ldos.sum() / vol
        dos_geom = math.exp(np.log(ldos).sum() / vol)
        tdos[wi] = (w.real, w.imag, dos_geom, 0.0, dos_val, 0.0, dos_geom / dos_val, 0.0)
        if save_only_dos:
            continue
        gr_re, gr_im = np.zeros((L0, L1)), np.zeros((L0, L1))
        np.add.at(gr_re, (d0, d1), g_re / vol)
        np.add.at(gr_im, (d0, d1), g_im / vol)
        gk = np.fft.fft2(gr_re + 1j * gr_im)
        per_w[_wstring(w)] = dict(tdos=tdos[wi].copy(), gr_full_re=g_re, gr_full_im=g_im, gr_re=gr_re, gr_im=gr_im, gk_re=gk.real, gk_im=gk.imag)
    return dict(tdos_gwr=tdos, per_w=per_w)


# ---- plaintext twin of the reference output (prog/data_save.hxx:9-30, prog/data_save.hpp:124-156, README.md:42-43) ----
def savetxt(fname, rows):
    """gftools-style plaintext: scientific notation, space separated, one row per line (README example:
    `5.000000e+01 -2.474323e-01 1.207943e-28 1.554312e-15`)."""
    rows = np.atleast_2d(np.asarray(rows, dtype=np.float64))
    with open(fname, "w") as fh:
        for r in rows:
            fh.write(" ".join("%e" % v for v in r) + "\n")


def save_binning_plaintext(outdir, name, rows):
    """save_binning(..., save_plaintext=true): <name>_binning.dat = nbins x 5 [n, mean, variance, stderr, tau_int] and, at the bin
    chosen by estimate_bin, <name>_error.dat = [n, mean, variance, stderr] (== /binning/<name> and /stats/<name> of the HDF5 file)."""
    import os
    cor = calc_cor_length(rows)
    table = [list(r) + [c] for r, c in zip(rows, cor)]
    savetxt(os.path.join(outdir, name + "_binning.dat"), table)
    b = estimate_bin(rows)
    savetxt(os.path.join(outdir, name + "_error.dat"), [list(rows[b])])
    return b


def save_energy_plaintext(outdir, energies, d2energies, beta, volume, max_depth=None):
    """The energy part of data_saver::save_all (prog/data_save.hxx:158-199) in the --plaintext layout."""
    rep = energy_report(energies, d2energies, beta, volume, max_depth)
    for name in ("energy", "d2energy", "cv"):
        save_binning_plaintext(outdir, name, rep[name]["binning"])
    return rep
