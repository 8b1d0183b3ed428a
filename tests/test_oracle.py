"""CPU tests: the oracle against the reference's own golden values, analytic spectra, LAPACK and the
committed fixtures (tests/golden).  No GPU needed."""
import math

import numpy as np
import pytest
import scipy.linalg as sl
import scipy.special

import oracle_lib as o


# ---------------- lattice (test/lattice_test.cpp:13-20, test/honeycomb_test.cpp:21-27) ----------------
def test_index_pos_roundtrip_and_order():
    L = 8
    for i in range(L * L):
        y, x = o.index_to_pos(o.CUBIC2D, L, i)[:2]
        assert y * L + x == i  # last coordinate fastest (hypercubic.cpp:31-51)
    z, y, x = o.index_to_pos(o.CUBIC3D, 4, 27)
    assert (z, y, x) == (1, 2, 3)


@pytest.mark.parametrize("kind,L,nnz_per_row", [(o.CUBIC1D, 8, 2), (o.CUBIC2D, 8, 4), (o.CUBIC3D, 4, 6), (o.TRIANGULAR, 6, 6),
                                                (o.HONEYCOMB, 8, 3), (o.HONEYCOMB_REF, 8, 3)])
def test_hopping_counts(kind, L, nnz_per_row):
    H = o.hopping_dense(kind, L)
    assert ((H != 0).sum(axis=1) == nnz_per_row).all()
    assert H.sum() == pytest.approx(-nnz_per_row * H.shape[0])


def test_honeycomb_sum_golden():
    # test/honeycomb_test.cpp:26: hopping_m().sum() == -3 L^2 for L = 8 (literal and intended)
    assert o.hopping_dense(o.HONEYCOMB_REF, 8).sum() == -3 * 64
    assert o.hopping_dense(o.HONEYCOMB, 8).sum() == -3 * 64


def test_honeycomb_literal_is_nonsymmetric_intended_is_symmetric():
    Hl = o.hopping_dense(o.HONEYCOMB_REF, 8)
    Hi = o.hopping_dense(o.HONEYCOMB, 8)
    assert np.abs(Hl - Hl.T).max() > 0
    assert np.abs(Hi - Hi.T).max() == 0
    assert np.abs(np.linalg.eigvalsh(Hi)).max() <= 3 + 1e-12
    Hm = o.hopping_dense(o.HONEYCOMB_REF_LOWER, 8)
    assert np.array_equal(np.tril(Hm), np.tril(Hl))
    assert np.abs(Hm - Hm.T).max() == 0


def test_L_below_3_rejected():
    with pytest.raises(RuntimeError):
        o.hopping_dense(o.CUBIC2D, 2)


# ---------------- E_ff (test/config_test.cpp:46-72) ----------------
def test_ff_energy_golden():
    L = 12
    W = [0, 1, 2]
    f = np.array([x % 2 for x in range(L)], dtype=np.int32)
    f[0] = 1
    assert o.ff_energy(f, W) == pytest.approx(2 * L / 2 * W[2] + 4 * W[1], abs=1e-15)


# ---------------- eigenvalues: analytic cases + LAPACK ----------------
def test_checkerboard_4x4_spectrum_and_logz():
    # test/config_test.cpp:10-42 (prints only); values derived in SURVEY section 4
    L = 4
    f = np.array([(x + y) % 2 for y in range(L) for x in range(L)], dtype=np.int32)
    r = o.calc_ed(o.CUBIC2D, L, f, 1.0, 0.5, 1.0, t=-1.0)
    ek = np.array([2 * (math.cos(2 * math.pi * a / L) + math.cos(2 * math.pi * b / L)) for a in range(L) for b in range(L)])
    exact = np.sort(np.concatenate([np.sqrt(ek[ek > 1e-9] ** 2 + 0.25), -np.sqrt(ek[ek > 1e-9] ** 2 + 0.25),
                                    [0.5] * 3, [-0.5] * 3]))
    assert np.abs(r["spectrum"] - exact).max() < 1e-13
    assert r["logZ"] == pytest.approx(17.615291438320348, rel=1e-14)


@pytest.mark.parametrize("kind,L,d", [(o.CUBIC1D, 12, 1), (o.CUBIC2D, 8, 2), (o.CUBIC3D, 4, 3)])
def test_free_lattice_spectrum(kind, L, d):
    n = L ** d
    mu = 0.3
    r = o.calc_ed(kind, L, np.zeros(n, np.int32), 1.0, mu, 2.0)
    ks = np.stack(np.meshgrid(*[np.arange(L)] * d, indexing="ij"), -1).reshape(-1, d)
    exact = np.sort(-2 * np.cos(2 * np.pi * ks / L).sum(axis=1) - mu)
    assert np.abs(r["spectrum"] - exact).max() < 1e-12


@pytest.mark.parametrize("n", [1, 2, 3, 17, 64, 200])
def test_eigh_against_lapack(n):
    rng = np.random.default_rng(n)
    A = rng.normal(size=(n, n))
    A = A + A.T
    ev, V = o.eigh(A, vectors=True)
    ref = sl.eigh(A, eigvals_only=True, driver="evd")
    scale = max(1.0, np.abs(ref).max())
    assert np.abs(ev - ref).max() <= 1e-12 * scale
    assert np.abs(A @ V - V * ev).max() <= 1e-11 * scale
    assert np.abs(V.T @ V - np.eye(n)).max() <= 1e-12


def test_eigh_reads_lower_triangle_only():
    rng = np.random.default_rng(0)
    A = rng.normal(size=(20, 20))
    S = np.tril(A) + np.tril(A, -1).T
    assert np.abs(o.eigh(np.tril(A) + 7 * np.triu(rng.normal(size=(20, 20)), 1)) - np.linalg.eigvalsh(S)).max() < 1e-12


def test_logz_formula_identity():
    # test/fast_update_test.cpp:78: shifted sum == direct sum(log(1+exp(-beta e))) to 1e-15 relative
    L, U, beta = 24, 8.0, 1 / 0.16
    f, _ = o.randomize_f(32167, L * L, L * L // 2)
    r = o.calc_ed(o.CUBIC2D, L, f, U, U / 2, beta, t=-1.0)
    s2 = float(np.sum(np.log(1 + np.exp(-beta * r["spectrum"]))))
    assert abs(r["logZ"] - s2) / abs(s2) <= 1e-15 * 4
    assert np.allclose(r["cached_fermi"], 1 / (1 + np.exp(beta * r["spectrum"])), rtol=1e-14)


def test_spectra_fixtures(golden):
    for c in golden["spectra"]["cases"]:
        r = o.calc_ed(o.KINDS[c["kind"]], c["L"], np.array(c["f"], np.int32), c["U"], c["mu_c"], c["beta"])
        ref = np.array(c["spectrum"])
        assert np.abs(r["spectrum"] - ref).max() <= 1e-10 * np.abs(ref).max()
        assert abs(r["logZ"] - c["logZ"]) <= 1e-10 * max(1, abs(c["logZ"]))


# ---------------- Chebyshev evaluator (test/fast_update_test.cpp:62-65) ----------------
def test_chebyshev_quadrature_goldens():
    M = G = 12  # cheb_size = 2*int(ln 576), grid = cheb_size in the reference test
    T, x, th = o.cheb_table(M, G)
    assert o.cheb_moment(M, G, np.ones(G), 0) == pytest.approx(1.0, abs=1e-8)
    assert o.cheb_moment(M, G, x * x, 0) == pytest.approx(0.5, abs=1e-8)
    assert o.cheb_moment(M, G, np.sin(x), 1) == pytest.approx(0.440051, abs=1e-6)
    assert o.cheb_moment(M, G, np.sin(x), 5) == pytest.approx(0.000249758, abs=1e-9)
    assert o.cheb_moment(M, G, np.sin(x), 1) == pytest.approx(scipy.special.jv(1, 1.0), abs=1e-12)
    assert o.cheb_moment(M, G, np.sin(x), 5) == pytest.approx(scipy.special.jv(5, 1.0), abs=1e-12)


def test_cheb_sizes():
    # fk_mc.hxx:60-63 at the BASELINE configs
    assert o.cheb_sizes(64, 2.2) == (10, 20)
    assert o.cheb_sizes(256, 2.2) == (12, 24)
    assert o.cheb_sizes(512, 2.2) == (14, 28)
    assert o.cheb_sizes(576, 2.2) == (14, 28)
    assert o.cheb_sizes(1024, 2.2) == (16, 32)


def test_kpm_vs_ed_reference_tolerance():
    # test/fast_update_test.cpp:29-77: 24x24, U=8, T=0.16, seed 32167, M = G = 12: |logZ_cheb - logZ_ed|/|logZ| <= 5e-2
    L, U, beta = 24, 8.0, 1 / 0.16
    f, _ = o.randomize_f(32167, L * L, L * L // 2)
    ed = o.calc_ed(o.CUBIC2D, L, f, U, U / 2, beta, t=-1.0)
    c = o.calc_chebyshev(o.CUBIC2D, L, f, U, U / 2, beta, 12, 12, t=-1.0)
    assert abs((c["logZ"] - ed["logZ"]) / c["logZ"]) <= 5e-2
    assert c["moments"][0] == 1.0


def test_kpm_fixtures_and_lanczos(golden):
    for c in golden["spectra"]["cases"]:
        if "kpm" not in c:
            continue
        f = np.array(c["f"], np.int32)
        k = o.KINDS[c["kind"]]
        r0 = o.calc_chebyshev(k, c["L"], f, c["U"], c["mu_c"], c["beta"], c["M"], c["G"], emode=0)
        assert np.abs(r0["moments"] - np.array(c["kpm"]["moments"])).max() <= 1e-11
        assert abs(r0["logZ"] - c["kpm"]["logZ"]) <= 1e-10 * abs(c["kpm"]["logZ"])
        r1 = o.calc_chebyshev(k, c["L"], f, c["U"], c["mu_c"], c["beta"], c["M"], c["G"], emode=1)
        assert abs(r1["e_min"] - c["kpm"]["e_min"]) <= 1e-11 and abs(r1["e_max"] - c["kpm"]["e_max"]) <= 1e-11
        assert abs(r1["logZ"] - c["kpm"]["logZ"]) <= 1e-10 * abs(c["kpm"]["logZ"])


def test_kpm_pruning_changes_little():
    f, _ = o.randomize_f(3, 256, 128)
    a = o.calc_chebyshev(o.CUBIC2D, 16, f, 2.0, 1.0, 10.0, 12, 24, prune=False)
    b = o.calc_chebyshev(o.CUBIC2D, 16, f, 2.0, 1.0, 10.0, 12, 24, prune=True)
    assert abs(a["logZ"] - b["logZ"]) <= 1e-9 * abs(a["logZ"])


# ---------------- RNG / Monte Carlo ----------------
def test_rng_fixtures(golden):
    g = golden["rng"]
    # std::mt19937's 10000th output for the default seed 5489 is 4123659995 (C++ standard [rand.predef])
    assert int(o.rng_stream(5489, 0, 0, 10000)[-1]) == 4123659995
    assert [int(x) for x in o.rng_stream(g["seed"], 0, 0, 40)] == g["raw"]
    assert [int(x) for x in o.rng_stream(g["seed"], 1, 64, 40)] == g["uniform_int_64"]
    # power-of-two range: site = top log2(V) bits of one word
    assert g["uniform_int_64"] == [w >> 26 for w in g["raw"]]
    u = o.rng_stream(g["seed"], 2, 0, 20)
    assert [float(x) for x in u] == g["uniform_real"][:20]
    assert u[0] == (g["raw"][0] + g["raw"][1] * 2.0 ** 32) / 2.0 ** 64
    f, _ = o.randomize_f(g["seed"], 64, 32)
    assert f.tolist() == g["randomize_f_8x8"] and f.sum() == 32


def test_mc_trace_fixture_and_consistency(golden):
    for t in golden["mc_trace"]["traces"]:
        p = o.make_params(kind=o.CUBIC2D, L=8, beta=t["beta"], U=t["U"], mc_flip=t["mc_flip"], mc_reshuffle=t["mc_reshuffle"],
                          cheb_moves=t["cheb"], seed=32167, nsweeps=3, sweep_len=16, ntherm_sweeps=1)
        r = o.mc_run(p, rank=t["rank"])
        tr = r["trace"]
        assert tr["accepted"].tolist() == t["accepted"] and tr["site_a"].tolist() == t["site_a"]
        assert np.allclose(tr["weight"], t["weight"], rtol=1e-9, atol=1e-12)
        assert np.allclose(r["energies"], t["energies"], rtol=1e-10)
        assert ((np.abs(tr["weight"]) > tr["u"]).astype(int) == tr["accepted"]).all()  # mc_metropolis.cpp:44
        assert r["naccept"] == tr["accepted"].sum()


def test_mc_addremove_weight_matches_direct_logz():
    p = o.make_params(kind=o.CUBIC2D, L=8, beta=4.0, U=4.0, nsweeps=1, sweep_len=8, ntherm_sweeps=0)
    r = o.mc_run(p)
    f, _ = o.randomize_f(32167, 64, 32)
    lz = o.calc_ed(o.CUBIC2D, 8, f, 4.0, 2.0, 4.0)["logZ"]
    tr = r["trace"]
    for s in range(8):
        g = f.copy()
        g[tr["site_a"][s]] ^= 1
        lz_new = o.calc_ed(o.CUBIC2D, 8, g, 4.0, 2.0, 4.0)["logZ"]
        w = math.exp(lz_new - lz) * (math.exp(4.0 * 2.0) if g[tr["site_a"][s]] else math.exp(-4.0 * 2.0))
        assert w == pytest.approx(tr["weight"][s], rel=1e-10)
        if tr["accepted"][s]:
            f, lz = g, lz_new
    assert (f == r["f_final"]).all()


def test_ipr_of_plane_waves():
    ed = o.calc_ed(o.CUBIC1D, 12, np.zeros(12, np.int32), 1.0, 0.0, 1.0, vectors=True)
    ipr = o.measure_ipr(ed["evecs"])
    assert ipr[0] == pytest.approx((12 * (1 / 12) ** 2) ** 0.25, rel=1e-12)  # uniform ground state


F5 = [0, 1, 0, 1, 1, 0, 1, 0, 0, 1, 0, 0, 1, 0, 1, 0, 1, 1, 1, 0, 1, 1, 0, 1, 1]
F7 = [0, 1, 0, 0, 0, 0, 1, 1, 0, 1, 1, 0, 0, 1, 1, 0, 1, 1, 0, 1, 1, 0, 1, 0, 0, 0, 1, 1, 0, 1, 1, 1, 1, 1, 0, 1, 1, 0, 1, 1, 1, 1, 0, 0, 0, 1, 0, 1, 1]


@pytest.mark.parametrize("U,f,golden", [(0.0, F5, 1.31597), (2.0, F5, 0.287547), (6.0, F5, 0.00260831), (0.37, F7, 1.26046)])
def test_stiffness_reference_goldens(U, f, golden):
    # test/stiffness_test.cpp:60-63: the reference's only known answers that go through eigenvalues AND eigenvectors (+-1e-3)
    L = int(round(len(f) ** 0.5))
    st, _ = o.stiffness(o.CUBIC2D, L, np.array(f, np.int32), U, U / 2, 1000.0)
    assert st == pytest.approx(golden, abs=1e-3)
    assert st == pytest.approx(golden, rel=2e-5)  # agrees to the printed digits
