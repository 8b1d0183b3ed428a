"""Host statistics (fk_mc_b200/stats.py) against the reference's Mathematica goldens (test/binning_test.cpp:60-63,
test/jackknife_test.cpp:56-134) and against the oracle restatement."""
import math

import numpy as np
import pytest

import oracle_lib as o
from fk_mc_b200 import stats

A = np.array([0.0711992, 0.344949, 0.940913, 0.166604, 0.811305, 0.617859, 0.462844, 0.550449, 0.28126, 0.0560575, 0.0673806, 0.710085,
              0.459742, 0.977218, 0.500193, 0.45763, 0.752903])                                   # test/binning_test.cpp:32-34
B = np.array([0.0203252, 0.0491541, 0.0942537, 0.0815104, 0.0569245, 0.0459406, 0.0963107, 0.0170473, 0.0730589, 0.012731, 0.0571613,
              0.0889872, 0.0771268, 0.0857561, 0.0130207, 0.0378117, 0.0690792])                 # test/jackknife_test.cpp:37-39


def test_binning_goldens():
    correct = [(17, 0.484035, 0.0872001, 0.07162), (8, 0.467231, 0.0422793, 0.0726974), (4, 0.467231, 0.0269458, 0.0820759),
               (2, 0.467231, 0.00162846, 0.0285348)]
    cor = [0, -0.0151463, 0.118023, -0.4253]
    for rows in (stats.accumulate_binning(A, 3), [tuple(r[:4]) for r in o.binning(A, 3)]):
        for got, want in zip(rows, correct):
            assert got[0] == want[0] and np.allclose(got[1:], want[1:], atol=1e-5)
    assert np.allclose(stats.calc_cor_length(stats.accumulate_binning(A, 3)), cor, atol=1e-5)
    assert np.allclose(o.binning(A, 3)[:, 4], cor, atol=1e-5)
    with pytest.raises(ValueError):
        stats.bin_series(A, 6)  # binning.hpp:66-68: step 64 > 17 samples


def test_jackknife_goldens():
    f1 = lambda x: x  # noqa: E731
    f2 = lambda e, e2, de2: e2 - de2 - e * e  # noqa: E731
    for depth, want in [(0, (17, 0.484035, 0.0872002, 0.07162)), (1, (8, 0.467231, 0.0422793, 0.0726974)),
                        (2, (4, 0.467231, 0.0269458, 0.0820759)), (3, (2, 0.467231, 0.00162846, 0.0285348))]:
        got = stats.jack(f1, [A], depth)
        assert got[0] == want[0] and np.allclose(got[1:], want[1:], atol=1e-5)
        assert np.allclose(o.jackknife(A, depth), want, atol=1e-5)
    for depth, want in [(0, (17, 0.0297767, 0.00815104, 0.0218969)), (1, (8, 0.0309895, 0.00214726, 0.0163832)),
                        (2, (4, 0.0324411, 0.00123011, 0.0175365))]:
        got = stats.jack(f2, [A, A * A, B], depth)
        assert got[0] == want[0] and np.allclose(got[1:], want[1:], atol=1e-5)
        assert np.allclose(o.jackknife(np.stack([A, A * A, B]), depth), want, atol=1e-5)
    rows = stats.accumulate_jackknife(f1, [A], 3)
    assert rows[3][0] == 2 and np.allclose(rows[3][1:], (0.467231, 0.00162846, 0.0285348), atol=1e-5)


def test_estimate_bin_and_energy_report():
    # the plateau search (data_save.hpp:108-122) walks down only while the relative change keeps shrinking and is < 5 %
    rows = [(64, 0, 1, 0.100), (32, 0, 1, 0.101), (16, 0, 1, 0.102), (8, 0, 1, 0.103)]
    assert stats.estimate_bin(rows) == 2
    rows = [(64, 0, 1, 0.103), (32, 0, 1, 0.1030), (16, 0, 1, 0.10301), (8, 0, 1, 0.104)]
    assert stats.estimate_bin(rows) == 0  # 0.96 % -> 0.0097 % -> 0 %: keeps shrinking all the way down
    rows = [(64, 0, 1, 0.05), (32, 0, 1, 0.08), (16, 0, 1, 0.100), (8, 0, 1, 0.101)]
    assert stats.estimate_bin(rows) == 2
    rng = np.random.default_rng(0)
    e = rng.normal(-50, 1, 4096)
    d2 = np.full(4096, 0.3)
    rep = stats.energy_report(e, d2, beta=2.0, volume=64)
    assert abs(rep["energy"]["stats"][1] + 50) < 5 * rep["energy"]["stats"][3]
    cv = rep["cv"]["stats"]
    assert abs(cv[1] - 4 * (1.0 - 0.3) / 64) < 5 * cv[3]


def test_plaintext_layout(tmp_path):
    # README.md:42-43: "<obs>_error.dat" holds 4 numbers in scientific notation, space separated
    rng = np.random.default_rng(1)
    e = rng.normal(-0.25 * 64, 0.5, 256)
    d2 = rng.normal(0.1, 0.01, 256)
    rep = stats.save_energy_plaintext(str(tmp_path), e, d2, beta=1.0, volume=64)
    err = open(tmp_path / "energy_error.dat").read().split()
    assert len(err) == 4 and all("e" in x for x in err)
    assert float(err[0]) == rep["energy"]["stats"][0] and float(err[1]) == pytest.approx(rep["energy"]["stats"][1], rel=1e-6)
    table = np.loadtxt(tmp_path / "cv_binning.dat")
    assert table.shape == (len(rep["cv"]["binning"]), 5)
    assert table[0, 0] == 256 and table[1, 0] == 128


def test_multi_chain_series_are_pooled_chain_major():
    """[measurement][chain] series must be binned along Monte Carlo time inside a chain (the reference gathers rank after rank,
    src/measures/energy.cpp:32-47): with strongly autocorrelated chains the error bar has to GROW with the bin level; a
    measurement-major flattening averages independent chains first and reports a flat (too small) error."""
    rng = np.random.default_rng(7)
    n_meas, n_chains, rho = 512, 64, 0.95
    x = np.zeros((n_meas, n_chains))
    x[0] = rng.normal(size=n_chains)
    for m in range(1, n_meas):
        x[m] = rho * x[m - 1] + math.sqrt(1 - rho * rho) * rng.normal(size=n_chains)   # AR(1), tau_int ~ 19
    pooled = stats.pool_chains(x)
    assert np.array_equal(pooled[:n_meas], x[:, 0]) and np.array_equal(pooled[n_meas:2 * n_meas], x[:, 1])
    assert np.array_equal(stats.pool_chains(pooled), pooled)
    with pytest.raises(ValueError):
        stats.pool_chains(np.zeros((2, 2, 2)))
    rep = stats.energy_report(x, x * x, 1.0, 1.0, max_depth=8)
    err = [r[3] for r in rep["energy"]["binning"]]
    assert err[6] > 3.0 * err[0]                      # autocorrelation visible: (1+rho)/(1-rho) = 39 -> factor ~ 6
    wrong = [r[3] for r in stats.accumulate_binning(x.reshape(-1)[::-1], 6)]
    assert wrong[5] < 1.5 * wrong[0]                  # what the measurement-major flattening used to report
    # true error of the mean of all samples, from the exact AR(1) autocorrelation time
    true_err = math.sqrt((1 + rho) / (1 - rho) / x.size)
    assert 0.6 * true_err < err[8] < 1.6 * true_err


def test_fstats_report():
    """f-sector statistics (prog/data_save.hxx:200-236) against direct formulas at bin level 0 and against the oracle's jackknife."""
    rng = np.random.default_rng(3)
    nf0 = rng.integers(20, 44, size=(64, 4)).astype(float)
    nfpi = np.abs(rng.integers(-10, 10, size=(64, 4))).astype(float)
    rep = stats.fstats_report(nf0, nfpi, max_depth=3)
    x = stats.pool_chains(nf0)
    assert rep["nf_0"]["binning"][0][1] == pytest.approx(x.mean(), rel=1e-14)
    n = x.size
    # jackknife of x2 - x^2 at depth 0: the bias-corrected value is the unbiased variance * (n - 1) / n ... check against the oracle instead
    ref = o.jackknife(np.stack([x, x * x, np.zeros_like(x)]), 0)   # oracle functor: e2 - de2 - e^2 with de2 = 0  ->  <x^2> - <x>^2
    got = rep["fsusc_0"]["binning"][0]
    assert got[1] == pytest.approx(ref[1], rel=1e-12) and got[3] == pytest.approx(ref[3], rel=1e-10)
    b = rep["binder_pi"]["stats"]
    y = stats.pool_chains(nfpi)
    assert b[1] == pytest.approx(1.0 - (y ** 4).mean() / 3.0 / (y ** 2).mean() ** 2, abs=5 * b[3] + 1e-12)


def test_dos_and_ipr_reports():
    """save_glocal / save_ipr post-processing (prog/data_save.hxx:265-345,487-532) against the defining formulas in complex arithmetic."""
    rng = np.random.default_rng(11)
    n_meas, n_chains, vol, off, beta = 32, 3, 16, 0.05, 4.0
    sp = np.sort(rng.normal(size=(n_meas, n_chains, vol)), axis=2)
    ip = rng.uniform(0.2, 1.0, size=(n_meas, n_chains, vol))
    wg = np.linspace(-2.0, 2.0, 9)
    rep = stats.dos_report(sp, wg, off, beta, max_depth=3)
    rows = np.transpose(sp, (1, 0, 2)).reshape(-1, vol)                       # chain after chain
    d0 = np.array([-(1.0 / (0.0 - r + 1j * off)).sum().imag / math.pi / vol for r in rows])
    assert rep["dos0"]["binning"][0][1] == pytest.approx(d0.mean(), rel=1e-12)
    assert rep["dos0"]["binning"][0][3] == pytest.approx(stats.calc_stats(d0)[3], rel=1e-12)
    b = rep["dos0"]["bin"]
    dw = np.array([-(1.0 / (wg[5] - r + 1j * off)).sum().imag / math.pi / vol for r in rows])
    st = stats.bin_stats(dw[::-1], b)
    assert rep["dos_err"][5] == pytest.approx((wg[5], st[1], st[3]), rel=1e-12)
    assert rep["dos_err"][:, 1].min() > 0 and rep["nc"][1] > 0 and rep["nc"][3] > 0
    ir = stats.ipr_report(sp, ip, wg, off, max_depth=3)
    rows_i = np.transpose(ip, (1, 0, 2)).reshape(-1, vol)
    g = [(1.0 / (0.0 - r + 1j * off)) for r in rows]
    i0 = np.array([(gi * ri ** 4).sum().imag / gi.sum().imag for gi, ri in zip(g, rows_i)])     # ipr_moment_f's nom / denom
    assert ir["ipr0"]["binning"][0][1] == pytest.approx(i0.mean(), rel=1e-12)
    assert ir["ipr_err"].shape == (9, 3) and (ir["ipr_err"][:, 1] > 0).all()
    # single-chain input [measurement][N] is accepted as is
    one = stats.dos_report(sp[:, 0], wg, off, beta, max_depth=2)
    assert one["dos0"]["binning"][0][0] == n_meas


def test_gwr_report_against_direct_inverse():
    """save_gwr (prog/data_save.hxx:535-709): G(w) from eigenvectors equals (w + i xi - H)^-1 averaged over the configurations; the translation
    average and its DFT follow the reference's index conventions; dataset names use the reference's trimmed std::to_string(float)."""
    L, U, xi = 4, 2.0, 0.05
    n = L * L
    H0 = o.hopping_dense(o.CUBIC2D, L)
    rng = np.random.default_rng(5)
    fs = (rng.random((3, n)) < 0.5).astype(float)
    evs, sps = [], []
    for f in fs:
        ev, V = np.linalg.eigh(H0 + np.diag(U * f - U / 2))
        evs.append(V)
        sps.append(ev)
    rep = stats.gwr_report(np.array(evs), np.array(sps), [0.0, 0.5], xi, (L, L))
    assert set(rep["per_w"]) == {"0._0.05", "0.5_0.05"}
    for w0, key in ((0.0, "0._0.05"), (0.5, "0.5_0.05")):
        G = np.mean([np.linalg.inv((w0 + 1j * xi) * np.eye(n) - (H0 + np.diag(U * f - U / 2))) for f in fs], axis=0)
        r = rep["per_w"][key]
        assert np.abs(r["gr_full_re"] - G.real).max() < 1e-10 and np.abs(r["gr_full_im"] - G.imag).max() < 1e-10
        ldos = -np.diag(G.imag) / math.pi
        assert r["tdos"][4] == pytest.approx(ldos.mean(), rel=1e-10) and r["tdos"][2] == pytest.approx(math.exp(np.log(ldos).mean()), rel=1e-10)
        assert r["tdos"][6] <= 1.0 + 1e-12                       # typical <= average (AM-GM)
        # G(r1 - r2): element (dy, dx) averages G(i, j) over all i with r_j = r_i + (dy, dx)
        acc = 0.0
        for i in range(n):
            y, x = divmod(i, L)
            acc += G[i, ((y + 1) % L) * L + (x + 2) % L]
        assert r["gr_re"][1, 2] + 1j * r["gr_im"][1, 2] == pytest.approx(acc / n, rel=1e-10)
        gk = np.fft.fft2(r["gr_re"] + 1j * r["gr_im"])
        assert np.allclose(r["gk_re"], gk.real) and np.allclose(r["gk_im"], gk.imag)
        assert r["gk_re"][0, 0] + 1j * r["gk_im"][0, 0] == pytest.approx(G.sum() / n, rel=1e-10)   # k = 0 component
    only = stats.gwr_report(np.array(evs)[:, None], np.array(sps)[:, None], [0.0], xi, (L, L), save_only_dos=True)   # [measurement][chain] input
    assert only["per_w"] == {} and only["tdos_gwr"][0, 4] == pytest.approx(rep["tdos_gwr"][0, 4], rel=1e-12)


def test_fcorrel_report():
    """save_fcorrel (prog/data_save.hxx:347-420) against a loop transcription of its formula on a small lattice."""
    rng = np.random.default_rng(8)
    L, n_meas = 4, 32
    V = L * L
    # checkerboard-biased occupations so that the correlator has structure
    idx = np.arange(V)
    stag = ((idx // L + idx % L) % 2 == 0)
    fo = (rng.random((n_meas, V)) < np.where(stag, 0.8, 0.2)).astype(np.int32)
    rep = stats.fcorrel_report(fo, (L, L), max_depth=2)
    b = rep["bin"]
    series = [fo[::-1, i].astype(float) for i in range(V)]
    nfb = stats.estimate_bin(stats.accumulate_binning(series[0], 2))
    nf_mean = np.array([stats.bin_stats(x, nfb)[1] for x in series])

    def f_loops(means, l):
        out = 0.0
        for i in range(V):
            y, x = divmod(i, L)
            for (yl, xl), (yr, xr) in ((((y - l) % L, x), ((y + l) % L, x)), ((y, (x - l) % L), (y, (x + l) % L))):
                out += (means[i] - nf_mean[i]) * (means[yl * L + xl] - nf_mean[yl * L + xl])
                out += (means[i] - nf_mean[i]) * (means[yr * L + xr] - nf_mean[yr * L + xr])
        return out / V / 4.0

    for l in (0, 1, 2):
        ref = stats.jack(lambda *m: f_loops(m, l), series, b)
        assert rep["fcorrel"][l, 1] == pytest.approx(ref[1], rel=1e-10, abs=1e-14) and rep["fcorrel"][l, 2] == pytest.approx(ref[3], rel=1e-10, abs=1e-14)
    assert rep["fcorrel"].shape == (3, 5) and rep["fcorrel"][0, 3] == pytest.approx(1.0)
    assert rep["fcorrel"][1, 1] * rep["fcorrel"][0, 1] < 0          # nearest neighbours anticorrelated in a checkerboard-biased ensemble
    assert rep["fcorrel_q"].shape == (L,) and abs(rep["fcorrel_q"][L // 2].real) > abs(rep["fcorrel_q"][0].real)   # weight at q = pi
