"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle and the
committed golden fixtures.  Tolerances (north_star): eigenvalues and free energies within 1e-10
relative in FP64 -- |d eps| <= 1e-10 max|eps|, |d logZ| <= 1e-10 max(1,|logZ|);
RNG streams, sites and accept/reject sequences bit-exact."""
import math

import numpy as np
import pytest
import scipy.linalg as sl

import fk_mc_b200 as fk
import oracle_lib as o

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def ctx8():
    c = fk.Context("cubic2d", 8, max_batch=64)
    yield c
    c.close()


# ---------------- RNG ----------------
@pytest.mark.parametrize("mode,V", [(0, 0), (1, 64), (1, 256), (1, 576), (1, 1000), (2, 0)])
def test_device_rng_matches_libstdcxx(ctx8, mode, V):
    assert np.array_equal(ctx8.rng_stream(32167, mode, V, 4000), o.rng_stream(32167, mode, V, 4000))


def test_device_rng_fixture_and_negative_seed(ctx8, golden):
    g = golden["rng"]
    assert [int(x) for x in ctx8.rng_stream(g["seed"], 0, 0, 40)] == g["raw"]
    assert [int(x) for x in ctx8.rng_stream(g["seed"], 1, 576, 40)] == g["uniform_int_576"]
    assert [float(x) for x in ctx8.rng_stream(g["seed"], 2, 0, 40)] == g["uniform_real"]
    assert int(ctx8.rng_stream(5489, 0, 0, 10000)[-1]) == 4123659995
    assert np.array_equal(ctx8.rng_stream(-7, 0, 0, 16), o.rng_stream(-7, 0, 0, 16))


# ---------------- lattices ----------------
@pytest.mark.parametrize("kind,L", [("cubic1d", 8), ("cubic2d", 8), ("cubic3d", 4), ("triangular", 6), ("honeycomb", 8),
                                    ("honeycomb_ref_lower", 8), ("cubic2d", 3)])
def test_hopping_matrices(kind, L):
    c = fk.Context(kind, L)
    assert np.array_equal(c.hopping_dense(), o.hopping_dense(o.KINDS[kind], L))
    c.close()


def test_bad_lattices_rejected():
    with pytest.raises(fk.FkmcError):
        fk.Context("cubic2d", 2)
    with pytest.raises(fk.FkmcError):
        fk.Context("honeycomb", 7)  # "Need even size", hypercubic.cpp:182


# ---------------- stage level: tridiagonal eigenvalues, tridiagonalisation ----------------
@pytest.mark.parametrize("n", [2, 5, 64, 100, 256, 577, 1024])
def test_tridiag_bisection(ctx8, n):
    rng = np.random.default_rng(n)
    d = rng.normal(size=(4, n)) * 2
    e = rng.normal(size=(4, n - 1))
    e[1, ::5] = 0.0      # decoupled blocks
    e[2] *= 1e-9         # nearly diagonal
    d[3] = 1.0           # constant diagonal ...
    e[3] = 0.0           # ... identity: fully degenerate
    ev = ctx8.tridiag_eigvals(d, e)
    for b in range(4):
        ref = sl.eigvalsh_tridiagonal(d[b], e[b]) if n > 2 else np.linalg.eigvalsh(np.diag(d[b]) + np.diag(e[b], 1) + np.diag(e[b], -1))
        assert np.abs(ev[b] - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())
        assert (np.diff(ev[b]) >= 0).all()


@pytest.mark.parametrize("n", [97, 1024])
def test_tridiag_tiny_offdiagonals(ctx8, n):
    """Runs of tiny off-diagonal entries shrink the Sturm pair (p_i, p_{i-1}) by e_i^2 per row: through the rescaling threshold, the
    denormal range and to zero inside one group of rows - the grouped recurrences must notice and take the guarded form."""
    rng = np.random.default_rng(n + 1)
    d = rng.normal(size=(6, n)) * 2
    e = rng.normal(size=(6, n - 1))
    e[0, 10:18] = 1e-40                                   # 8 in a row: 2^-2126, exact zero inside a group
    e[1, 5::7] *= 1e-30                                   # isolated tiny entries
    e[2, 20:24] = 3e-39                                   # 4 in a row: 2^-1020, the edge of the denormal range
    e[2, 40:45] = 1e-38
    e[3] *= 10.0 ** -rng.integers(0, 60, size=n - 1)      # every scale at once
    e[4, :] = 1e-25                                        # diagonal matrix to working precision
    d[5] = np.round(d[5])                                  # degenerate diagonal ...
    e[5] *= 1e-20                                          # ... barely coupled
    ev = ctx8.tridiag_eigvals(d, e)
    for b in range(6):
        ref = sl.eigvalsh_tridiagonal(d[b], e[b])
        assert np.abs(ev[b] - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max()), b
        assert (np.diff(ev[b]) >= 0).all()


def test_tridiag_matches_oracle_ql(ctx8):
    A = o.hopping_dense(o.CUBIC2D, 8) + np.diag(np.arange(64) % 2 * 1.0)
    d, s = o.tridiag(A)
    assert np.abs(ctx8.tridiag_eigvals(d, s)[0] - o.tridiag_eig(d, s)).max() <= 1e-13


def test_tridiag_large_scale_entries(ctx8):
    rng = np.random.default_rng(5)
    d, e = rng.normal(size=(1, 300)) * 1e12, rng.normal(size=(1, 299)) * 1e12
    ref = sl.eigvalsh_tridiagonal(d[0], e[0])
    assert np.abs(ctx8.tridiag_eigvals(d, e)[0] - ref).max() <= 1e-13 * np.abs(ref).max()


@pytest.mark.parametrize("n", [2, 3, 31, 32, 33, 64, 100, 257])
def test_sytrd_preserves_spectrum(ctx8, n):
    rng = np.random.default_rng(n)
    A = rng.normal(size=(3, n, n))
    A = A + np.transpose(A, (0, 2, 1))
    A[2] = np.diag(rng.normal(size=n))  # already diagonal: every reflector is trivial (tau = 0)
    lower = np.tril(A)  # the kernel must read the lower triangle only
    d, e = ctx8.sytrd(lower + 5.0 * np.triu(np.ones((n, n)), 1))
    for b in range(3):
        ref = sl.eigvalsh(A[b])
        got = sl.eigvalsh_tridiagonal(d[b], e[b]) if n > 2 else np.linalg.eigvalsh(np.array([[d[b, 0], e[b, 0]], [e[b, 0], d[b, 1]]]))
        assert np.abs(ref - got).max() <= 1e-12 * max(1.0, np.abs(ref).max())
        # similarity invariants: trace and Frobenius norm
        assert abs(d[b].sum() - np.trace(A[b])) <= 1e-11 * max(1.0, np.abs(ref).max()) * n
        assert abs((d[b] ** 2).sum() + 2 * (e[b] ** 2).sum() - (A[b] ** 2).sum()) <= 1e-10 * (A[b] ** 2).sum() + 1e-12


def _band_eigvals(AB):
    # AB[d][c] = A(c+d, c): scipy's lower band form
    return sl.eig_banded(AB, lower=True, eigvals_only=True)


@pytest.mark.parametrize("n", [8, 16, 33, 100, 256, 512, 520, 576, 1024])
def test_two_stage_reduction_preserves_spectrum(ctx8, n):
    """dense -> band (sy2sb, half-bandwidth 8) -> tridiagonal (sb2st) against LAPACK; n >= 512 runs the tiled look-ahead kernel."""
    rng = np.random.default_rng(n)
    A = rng.normal(size=(3, n, n))
    A = A + np.transpose(A, (0, 2, 1))
    A[2] = np.diag(rng.normal(size=n)) + np.diag(np.ones(n - 1), 1) + np.diag(np.ones(n - 1), -1)  # already tridiagonal: trivial reflectors
    AB = ctx8.sy2sb(np.tril(A) + 7.0 * np.triu(np.ones((n, n)), 1))  # only the lower triangle may be read
    d, e = ctx8.sb2st(AB)
    for b in range(3):
        ref = sl.eigvalsh(A[b])
        scale = max(1.0, np.abs(ref).max())
        assert np.abs(_band_eigvals(AB[b]) - ref).max() <= 1e-12 * scale
        assert np.abs(sl.eigvalsh_tridiagonal(d[b], e[b]) - ref).max() <= 1e-12 * scale
        assert abs(AB[b, 0].sum() - np.trace(A[b])) <= 1e-11 * scale * n
        fro = (AB[b, 0] ** 2).sum() + 2 * (AB[b, 1:] ** 2).sum()
        assert abs(fro - (A[b] ** 2).sum()) <= 1e-10 * (A[b] ** 2).sum()


def test_two_stage_full_size_batch():
    """BASELINE config 5's measurement eigensolve at more than two waves of CTAs (the per-SM scratch slots of sy2sb are reused):
    size-independent invariants for every matrix, oracle spot checks for a few."""
    B, L, U, beta = 333, 32, 2.0, 20.0
    c = fk.Context("cubic2d", L, max_batch=B)
    n = c.N
    rng = np.random.default_rng(5)
    fs = (rng.random((B, n)) < 0.5).astype(np.int32)
    ev = c.logz_ed(fs, U, U / 2, beta)["spectrum"]
    assert (np.diff(ev, axis=1) >= 0).all()
    nf = fs.sum(axis=1)
    assert np.abs(ev.sum(axis=1) - (U * nf - U / 2 * n)).max() < 1e-9
    assert np.abs((ev ** 2).sum(axis=1) - (4 * n + ((U * fs - U / 2) ** 2).sum(axis=1))).max() < 1e-8
    r2 = c.logz_ed(1 - fs[:4], U, U / 2, beta)  # particle-hole symmetry on the bipartite lattice
    assert np.abs(r2["spectrum"] + ev[:4, ::-1]).max() < 1e-11
    for b in (0, 147, 148, 332):
        ref = o.calc_ed(o.CUBIC2D, L, fs[b], U, U / 2, beta)
        assert np.abs(ref["spectrum"] - ev[b]).max() <= TOL * np.abs(ref["spectrum"]).max()
    # bit-reproducible: the shared-memory accumulation order of the tile pass is fixed
    ev2 = c.logz_ed(fs[:160], U, U / 2, beta)["spectrum"]
    assert np.array_equal(ev2, ev[:160])
    c.close()


# ---------------- calc_ed ----------------
ED_CASES = [("cubic2d", 8, 1.0, 1.0), ("cubic2d", 16, 2.0, 10.0), ("cubic3d", 8, 4.0, 5.0), ("triangular", 24, 2.0, 10.0),
            ("honeycomb", 24, 2.0, 10.0), ("honeycomb_ref_lower", 24, 2.0, 10.0), ("cubic2d", 32, 2.0, 20.0), ("cubic1d", 12, 1.5, 3.0),
            ("cubic2d", 5, 0.37, 1000.0)]


@pytest.mark.parametrize("kind,L,U,beta", ED_CASES)
def test_calc_ed_matches_oracle(kind, L, U, beta):
    c = fk.Context(kind, L, max_batch=5)
    n = c.N
    fs = np.stack([o.randomize_f(32167 + i, n, n // 2)[0] for i in range(3)] + [np.zeros(n, np.int32), np.ones(n, np.int32)])
    r = c.logz_ed(fs, U, U / 2, beta, want_caches=True)
    for b in range(5):
        ref = o.calc_ed(o.KINDS[kind], L, fs[b], U, U / 2, beta)
        assert np.abs(ref["spectrum"] - r["spectrum"][b]).max() <= TOL * np.abs(ref["spectrum"]).max()
        if np.isfinite(ref["logZ"]):
            assert abs(ref["logZ"] - r["logZ"][b]) <= TOL * max(1.0, abs(ref["logZ"]))
        else:  # beta = 1000: the reference's shifted formula underflows to -inf (configuration.cpp:235-242); so must we
            assert r["logZ"][b] == ref["logZ"]
        assert (np.diff(r["spectrum"][b]) >= 0).all()
        fin = np.isfinite(ref["cached_exp"])
        assert np.allclose(r["cached_exp"][b][fin], ref["cached_exp"][fin], rtol=1e-9 * max(1.0, beta))
        assert np.allclose(r["cached_fermi"][b], ref["cached_fermi"], rtol=1e-9 * max(1.0, beta), atol=1e-300)
    c.close()


def test_calc_ed_golden_fixtures(golden):
    for cse in golden["spectra"]["cases"]:
        c = fk.Context(cse["kind"], cse["L"])
        r = c.logz_ed(np.array(cse["f"], np.int32), cse["U"], cse["mu_c"], cse["beta"])
        ref = np.array(cse["spectrum"])
        assert np.abs(r["spectrum"][0] - ref).max() <= TOL * np.abs(ref).max()
        assert abs(r["logZ"][0] - cse["logZ"]) <= TOL * max(1.0, abs(cse["logZ"]))
        c.close()


def test_calc_ed_analytic_cases():
    # checkerboard 4x4 (test/config_test.cpp:10-42, hopping +1) and its logZ(beta = 1)
    c = fk.Context("cubic2d", 4, t=-1.0)
    f = np.array([(x + y) % 2 for y in range(4) for x in range(4)], dtype=np.int32)
    r = c.logz_ed(f, 1.0, 0.5, 1.0)
    assert r["logZ"][0] == pytest.approx(17.615291438320348, rel=1e-13)
    assert np.abs(np.abs(r["spectrum"][0][[0, 15]]) - 4.031128874149275).max() < 1e-13
    c.close()
    # free cubic3d: eps = -2 sum cos k - mu (highly degenerate spectrum)
    c = fk.Context("cubic3d", 4)
    r = c.logz_ed(np.zeros(64, np.int32), 1.0, 0.3, 2.0)
    ks = np.stack(np.meshgrid(*[np.arange(4)] * 3, indexing="ij"), -1).reshape(-1, 3)
    exact = np.sort(-2 * np.cos(2 * np.pi * ks / 4).sum(axis=1) - 0.3)
    assert np.abs(r["spectrum"][0] - exact).max() < 1e-12
    c.close()


def test_calc_ed_full_size_properties():
    # BASELINE config 2 at full batch: cubic2d L=16, beta=10, U=2, 4096 configurations in one call
    B, L, U, beta = 4096, 16, 2.0, 10.0
    c = fk.Context("cubic2d", L, max_batch=B)
    rng = np.random.default_rng(0)
    fs = (rng.random((B, 256)) < 0.5).astype(np.int32)
    r = c.logz_ed(fs, U, U / 2, beta)
    ev = r["spectrum"]
    assert (np.diff(ev, axis=1) >= 0).all()
    # trace and Frobenius invariants of H = T + diag(U f - mu): sum eps = U nf - mu N, sum eps^2 = 4 N t^2 + sum diag^2
    nf = fs.sum(axis=1)
    assert np.abs(ev.sum(axis=1) - (U * nf - U / 2 * 256)).max() < 1e-9
    assert np.abs((ev ** 2).sum(axis=1) - (4 * 256 + ((U * fs - U / 2) ** 2).sum(axis=1))).max() < 1e-8
    direct = np.log1p(np.exp(-beta * ev)).sum(axis=1)
    assert np.abs(r["logZ"] - direct).max() <= TOL * np.abs(direct).max()
    # particle-hole symmetry at half filling: spectrum(f) = -spectrum(1-f) reversed on a bipartite lattice
    r2 = c.logz_ed(1 - fs[:8], U, U / 2, beta)
    assert np.abs(r2["spectrum"] + ev[:8, ::-1]).max() < 1e-11
    # spot check against the oracle
    for b in (0, 1777, 4095):
        ref = o.calc_ed(o.CUBIC2D, L, fs[b], U, U / 2, beta)
        assert np.abs(ref["spectrum"] - ev[b]).max() <= TOL * np.abs(ref["spectrum"]).max()
    c.close()


def test_energy_from_spectrum(ctx8):
    f, _ = o.randomize_f(1, 64, 32)
    ed = o.calc_ed(o.CUBIC2D, 8, f, 1.0, 0.5, 3.0)
    out = ctx8.energy_from_spectrum(ed["spectrum"], 3.0)[0]
    ex = ed["cached_exp"]
    assert out[0] == pytest.approx(np.sum(ed["spectrum"] / (1 + ex)), rel=1e-12)
    assert out[1] == pytest.approx(0.5 * np.sum(ed["spectrum"] ** 2 / (1 + 0.5 * (ex + 1 / ex))), rel=1e-12)
    assert out[2] == pytest.approx(ed["logZ"], rel=1e-12)


def test_argument_errors(ctx8):
    with pytest.raises(fk.FkmcError) as e:
        ctx8.logz_ed(np.zeros((65, 64), np.int32), 1.0, 0.5, 1.0)  # B > max_batch
    assert e.value.code == 1
    with pytest.raises(fk.FkmcError):
        ctx8.logz_ed(np.zeros((1, 63), np.int32), 1.0, 0.5, 1.0)
    with pytest.raises(fk.FkmcError):
        ctx8.logz_kpm(np.zeros((1, 64), np.int32), 1.0, 0.5, 1.0, 11, 22)  # odd M: assert(cheb_size%2==0), configuration.cpp:114
    with pytest.raises(fk.FkmcError):
        ctx8.chain_run_sweeps(1)  # before chain_init
    with pytest.raises(fk.FkmcError):
        ctx8.chain_init(4, 1.0, 1.0, mc_add_remove=0.0)  # "No registered moves", mc_metropolis.cpp:35-38


@pytest.mark.parametrize("kind,L", [("cubic2d", 20), ("honeycomb", 20), ("cubic2d", 28)])
def test_calc_ed_partial_tiles(kind, L):
    """N = 400 / 784: sizes that are not multiples of the 32 x 32 tile, through the 4-warp (N <= 512) and the 8-warp variant of the
    tiled dense->band kernel; a batch larger than the number of SMs so that scratch slots are reused."""
    B = 300
    c = fk.Context(kind, L, max_batch=B)
    n = c.N
    rng = np.random.default_rng(5)
    fs = (rng.random((B, n)) < 0.5).astype(np.int32)
    r = c.logz_ed(fs, 2.0, 1.0, 5.0)
    for b in (0, 7, B - 1):
        ref = o.calc_ed(o.KINDS[kind], L, fs[b], 2.0, 1.0, 5.0)
        assert np.abs(r["spectrum"][b] - ref["spectrum"]).max() <= TOL * np.abs(ref["spectrum"]).max()
        assert abs(r["logZ"][b] - ref["logZ"]) <= TOL * abs(ref["logZ"])
    # every matrix of the batch: trace and Frobenius norm of H from the spectrum
    tr = r["spectrum"].sum(axis=1)
    assert np.abs(tr - (2.0 * fs.sum(axis=1) - 1.0 * n)).max() <= 1e-9 * n
    c.close()


# ---------------- calc_chebyshev ----------------
KPM_CASES = [("cubic2d", 8, 1.0, 1.0), ("cubic2d", 16, 2.0, 10.0), ("cubic3d", 8, 4.0, 5.0), ("triangular", 24, 2.0, 10.0),
             ("honeycomb", 24, 2.0, 10.0), ("cubic2d", 32, 2.0, 20.0), ("cubic1d", 12, 1.5, 3.0), ("honeycomb_ref_lower", 8, 2.0, 4.0)]


@pytest.mark.parametrize("kind,L,U,beta", KPM_CASES)
def test_calc_chebyshev_matches_oracle(kind, L, U, beta):
    c = fk.Context(kind, L, max_batch=4)
    n = c.N
    M, G = fk.cheb_sizes(n, 2.2)
    fs = np.stack([o.randomize_f(32167 + i, n, n // 2)[0] for i in range(3)] + [np.zeros(n, np.int32)])
    r = c.logz_kpm(fs, U, U / 2, beta, M, G)
    for b in range(4):
        ref = o.calc_chebyshev(o.KINDS[kind], L, fs[b], U, U / 2, beta, M, G, emode=0)
        scale = max(abs(ref["e_min"]), abs(ref["e_max"]))
        assert abs(r["e_min"][b] - ref["e_min"]) <= TOL * scale and abs(r["e_max"][b] - ref["e_max"]) <= TOL * scale
        assert np.abs(r["moments"][b] - ref["moments"]).max() <= TOL
        assert abs(r["logZ"][b] - ref["logZ"]) <= TOL * max(1.0, abs(ref["logZ"]))
        assert r["a"][b] == pytest.approx((r["e_max"][b] - r["e_min"][b]) / 2) and r["moments"][b][0] == 1.0
    c.close()


@pytest.mark.parametrize("kind,L,U,beta,M", [("cubic2d", 32, 2.0, 20.0, 16), ("cubic2d", 16, 2.0, 10.0, 12), ("cubic2d", 24, 4.0, 5.0, 4),
                                             ("cubic2d", 32, 2.0, 5.0, 20), ("triangular", 24, 2.0, 10.0, 14),
                                             ("honeycomb", 24, 2.0, 10.0, 14), ("honeycomb", 16, 4.0, 5.0, 12)])
def test_kpm_two_kernel_path_matches_single_kernel_and_oracle(kind, L, U, beta, M):
    """kpm2d.cu (strip Lanczos + Laguerre Ritz values, ring-ordered patch recursion) against kpm.cu on a batch with
    empty / full / random configurations, and against the oracle on two of them (configuration.cpp:94-205)."""
    B = 40
    c = fk.Context(kind, L, max_batch=B)
    n = c.N
    rng = np.random.default_rng(7)
    fs = (rng.random((B, n)) < rng.random((B, 1))).astype(np.int32)
    fs[0], fs[1] = 0, 1
    r2 = c.logz_kpm(fs, U, U / 2, beta, M, 2 * M)
    c.set_option("kpm_v1", 1)
    r1 = c.logz_kpm(fs, U, U / 2, beta, M, 2 * M)
    c.set_option("kpm_v1", 0)
    c.set_option("kpm_generic_schedule", 1)      # run-time slot schedule instead of the compile-time one (cubic2d)
    r3 = c.logz_kpm(fs, U, U / 2, beta, M, 2 * M)
    c.set_option("kpm_generic_schedule", 0)
    assert np.abs(r3["moments"] - r2["moments"]).max() <= 1e-12
    scale = np.maximum(np.abs(r1["e_min"]), np.abs(r1["e_max"]))
    assert (np.abs(r1["e_min"] - r2["e_min"]) <= 1e-12 * scale).all() and (np.abs(r1["e_max"] - r2["e_max"]) <= 1e-12 * scale).all()
    assert np.abs(r1["moments"] - r2["moments"]).max() <= 1e-12
    assert (np.abs(r1["logZ"] - r2["logZ"]) <= 1e-12 * np.maximum(1.0, np.abs(r1["logZ"]))).all()
    for b in (1, 5):
        ref = o.calc_chebyshev(o.KINDS[kind], L, fs[b], U, U / 2, beta, M, 2 * M, emode=0)
        assert np.abs(r2["moments"][b] - ref["moments"]).max() <= TOL
        assert abs(r2["logZ"][b] - ref["logZ"]) <= TOL * max(1.0, abs(ref["logZ"]))
    c.close()


def test_calc_chebyshev_golden_and_other_sizes(golden):
    for cse in golden["spectra"]["cases"]:
        if "kpm" not in cse:
            continue
        c = fk.Context(cse["kind"], cse["L"])
        f = np.array(cse["f"], np.int32)
        r = c.logz_kpm(f, cse["U"], cse["mu_c"], cse["beta"], cse["M"], cse["G"])
        assert np.abs(r["moments"][0] - np.array(cse["kpm"]["moments"])).max() <= TOL
        assert abs(r["logZ"][0] - cse["kpm"]["logZ"]) <= TOL * abs(cse["kpm"]["logZ"])
        for M, G in [(2, 10), (4, 10), (6, 12), (20, 40), (32, 64)]:  # every template instance family + grid sizes
            if cse["L"] > 12:
                break
            ref = o.calc_chebyshev(o.KINDS[cse["kind"]], cse["L"], f, cse["U"], cse["mu_c"], cse["beta"], M, G)
            rr = c.logz_kpm(f, cse["U"], cse["mu_c"], cse["beta"], M, G)
            assert np.abs(rr["moments"][0] - ref["moments"]).max() <= TOL
            assert abs(rr["logZ"][0] - ref["logZ"]) <= TOL * max(1.0, abs(ref["logZ"]))
        c.close()


def test_kpm_reference_benchmark_tolerances():
    # test/fast_update_test.cpp:77 (24x24, U=8, T=0.16, seed 32167, M=G=12): KPM vs ED within 5e-2,
    # and benchmark/fast_update.cpp:67-72 weight agreement within 6e-2 for an add_remove attempt (L=16, U=1, T=0.1)
    L, U, beta = 24, 8.0, 1 / 0.16
    c = fk.Context("cubic2d", L, t=-1.0)
    f, _ = o.randomize_f(32167, L * L, L * L // 2)
    ed = c.logz_ed(f, U, U / 2, beta)["logZ"][0]
    kp = c.logz_kpm(f, U, U / 2, beta, 12, 12)["logZ"][0]
    assert abs((kp - ed) / kp) <= 5e-2
    c.close()
    L, U, beta = 16, 1.0, 10.0
    c = fk.Context("cubic2d", L, t=-1.0, max_batch=2)
    f, _ = o.randomize_f(32167, 256, 128)
    g = f.copy()
    site = int(o.rng_stream(32167, 1, 256, 1)[0])
    g[site] ^= 1
    fac = math.exp(beta * U / 2) if g[site] else math.exp(-beta * U / 2)  # moves.cpp:65
    M = int(math.log(256) * 2.35)
    M += M % 2
    e = c.logz_ed(np.stack([f, g]), U, U / 2, beta)["logZ"]
    k = c.logz_kpm(np.stack([f, g]), U, U / 2, beta, M, 2 * M)["logZ"]
    assert abs(math.exp(e[1] - e[0]) * fac - math.exp(k[1] - k[0]) * fac) < 6e-2
    c.close()


# ---------------- Markov chains: identical accept/reject sequences ----------------
CHAIN_CASES = [("cubic2d", 8, 1.0, 1.0, False, 0.0, 0.0), ("cubic2d", 8, 4.0, 4.0, False, 0.5, 0.1), ("cubic2d", 8, 4.0, 4.0, True, 0.5, 0.1),
               ("cubic2d", 16, 2.0, 10.0, False, 0.0, 0.0), ("cubic3d", 4, 4.0, 5.0, False, 1.0, 0.0), ("triangular", 6, 2.0, 10.0, True, 0.0, 0.0),
               ("honeycomb", 6, 2.0, 10.0, False, 0.3, 0.0), ("cubic2d", 16, 2.0, 10.0, True, 0.0, 0.0)]


def _compare_chain(c, ch, kind, L, U, beta, cheb, flip, resh, nsw, sl_, tr, se, st, seed, W=(), rank0=0):
    """Chain `ch` of the context (== reference rank rank0 + ch) against the oracle.  Weight tolerance: w = exp(dlogZ) and dlogZ
    must agree within 1e-10 max(1, |logZ|) (north_star / SURVEY H5), i.e. |dw| <= w * 1e-10 |logZ|; never tighter than 1e-9."""
    p = o.make_params(kind=o.KINDS[kind], L=L, beta=beta, U=U, mc_flip=flip, mc_reshuffle=resh, cheb_moves=cheb, seed=seed,
                      nsweeps=nsw, sweep_len=sl_, ntherm_sweeps=1, W=W)
    r = o.mc_run(p, rank=rank0 + ch)
    t = r["trace"]
    wtol = max(1e-9, TOL * float(np.abs(t["logz_new"]).max()))
    assert np.array_equal(t["u"], tr["u"][:, ch])                       # same RNG stream, same consumption order
    assert np.array_equal(t["move"][t["site_a"] >= 0], tr["move"][:, ch][t["site_a"] >= 0])
    assert np.array_equal(t["site_a"], tr["site_a"][:, ch]) and np.array_equal(t["site_b"], tr["site_b"][:, ch])
    w, wg = t["weight"], tr["weight"][:, ch]
    assert (np.abs(w - wg) <= wtol * np.maximum(1.0, np.abs(w))).all()     # ratio = exp(dlogZ): dlogZ within 1e-10 |logZ|
    # accept/reject identical except where |w| sits within tolerance of u (north_star); none expected at these sizes
    near = np.abs(np.abs(w) - t["u"]) <= wtol * np.maximum(1.0, np.abs(w))
    assert np.array_equal(t["accepted"][~near], tr["accepted"][:, ch][~near])
    assert near.sum() == 0
    assert np.array_equal(r["f_final"], st["f"][ch]) and r["naccept"] == st["naccept"][ch]
    assert np.abs(r["energies"] - se["energies"][:, ch]).max() <= 1e-9 * max(1.0, np.abs(r["energies"]).max())
    assert np.abs(r["d2energies"] - se["d2energies"][:, ch]).max() <= 1e-9 * max(1.0, np.abs(r["d2energies"]).max())
    assert np.abs(r["c_energies"] - se["c_energies"][:, ch]).max() <= 1e-9 * max(1.0, np.abs(r["c_energies"]).max())


@pytest.mark.parametrize("kind,L,U,beta,cheb,flip,resh", CHAIN_CASES)
def test_chain_matches_oracle(kind, L, U, beta, cheb, flip, resh):
    nch, nsw, sl_ = 4, 3, 16
    c = fk.Context(kind, L, max_batch=nch)
    c.chain_init(nch, beta, U, mc_flip=flip, mc_reshuffle=resh, cheb_moves=cheb, seed=32167, sweep_len=sl_, ntherm_sweeps=1,
                 measure_energy=True, record_trace=True, max_sweeps=nsw + 1)
    c.chain_run_sweeps(2)
    c.chain_run_sweeps(nsw - 1)  # resumable in pieces
    tr, se, st = c.chain_get_trace(), c.chain_get_series(), c.chain_get_state()
    assert tr["n_steps"] == (nsw + 1) * sl_ and se["n_measured"] == nsw
    for ch in range(nch):
        _compare_chain(c, ch, kind, L, U, beta, cheb, flip, resh, nsw, sl_, tr, se, st, 32167)
    c.close()


# The BASELINE sizes run kernels the small cases above never reach (tiled sy2sb<.,8> / <.,4>, the compile-time-schedule moments
# kernel at M/2 = 8, the 3-D generic KPM kernel): sequence identity is tested there too, on a few chains and sweeps.
BASELINE_CHAIN_CASES = [("cubic2d", 32, 2.0, 20.0, True, 2, 2),       # c5: KPM moves M=16 G=32 + exact measurement solve (N = 1024)
                        ("cubic2d", 32, 2.0, 20.0, False, 2, 1),      # exact moves at N = 1024 (sy2sb<., 8>)
                        ("cubic3d", 8, 4.0, 5.0, False, 3, 2),        # c3 (N = 512, sy2sb<., 4>)
                        ("cubic3d", 8, 4.0, 5.0, True, 2, 2),         # 3-D KPM (generic kernel)
                        ("triangular", 24, 2.0, 10.0, False, 3, 2),   # c4 (N = 576)
                        ("honeycomb", 24, 2.0, 10.0, False, 3, 2),
                        ("triangular", 24, 2.0, 10.0, True, 2, 2),
                        ("honeycomb", 24, 2.0, 10.0, True, 2, 2)]


@pytest.mark.parametrize("kind,L,U,beta,cheb,nch,nsw", BASELINE_CHAIN_CASES)
def test_chain_matches_oracle_baseline_sizes(kind, L, U, beta, cheb, nch, nsw):
    sl_ = 16
    c = fk.Context(kind, L, max_batch=nch)
    c.chain_init(nch, beta, U, cheb_moves=cheb, seed=32167, sweep_len=sl_, ntherm_sweeps=1, measure_energy=True, record_trace=True,
                 max_sweeps=nsw + 1)
    c.chain_run_sweeps(nsw + 1)
    tr, se, st = c.chain_get_trace(), c.chain_get_series(), c.chain_get_state()
    assert tr["n_steps"] == (nsw + 1) * sl_ and se["n_measured"] == nsw
    for ch in range(nch):
        _compare_chain(c, ch, kind, L, U, beta, cheb, 0.0, 0.0, nsw, sl_, tr, se, st, 32167)
    c.close()


def test_chain_out_of_full_c2_batch():
    """BASELINE config 2: 4096 chains of cubic2d L=16 batched on one GPU; chains picked across the batch reproduce their reference rank."""
    nch, nsw, sl_, L, U, beta = 4096, 1, 16, 16, 2.0, 10.0
    c = fk.Context("cubic2d", L, max_batch=nch)
    c.chain_init(nch, beta, U, seed=32167, sweep_len=sl_, ntherm_sweeps=1, record_trace=True, max_sweeps=nsw + 1)
    c.chain_run_sweeps(nsw + 1)
    tr, se, st = c.chain_get_trace(), c.chain_get_series(), c.chain_get_state()
    for ch in (0, 1777, 4095):
        _compare_chain(c, ch, "cubic2d", L, U, beta, False, 0.0, 0.0, nsw, sl_, tr, se, st, 32167)
    c.close()


@pytest.mark.parametrize("cheb", [False, True])
def test_chain_1d_with_ff_interaction(cheb):
    """1-D lattice with W != {}: exp(-beta dE_ff) in the weights and E_ff in the energy (src/moves.cpp:44-65, src/measures/energy.cpp:19,
    src/configuration.cpp:59-77); all three move kinds."""
    nch, nsw, sl_, L, U, beta, W = 3, 3, 16, 16, 2.0, 3.0, (0.3, -0.2, 0.15)
    c = fk.Context("cubic1d", L, max_batch=nch)
    c.chain_init(nch, beta, U, mc_flip=0.5, mc_reshuffle=0.2, cheb_moves=cheb, seed=99, sweep_len=sl_, ntherm_sweeps=1, record_trace=True,
                 max_sweeps=nsw + 1, W=W)
    c.chain_run_sweeps(nsw + 1)
    tr, se, st = c.chain_get_trace(), c.chain_get_series(), c.chain_get_state()
    for ch in range(nch):
        _compare_chain(c, ch, "cubic1d", L, U, beta, cheb, 0.5, 0.2, nsw, sl_, tr, se, st, 99, W=W)
    # the interaction really enters: without W the same seeds give another trajectory
    c.chain_init(nch, beta, U, mc_flip=0.5, mc_reshuffle=0.2, cheb_moves=cheb, seed=99, sweep_len=sl_, ntherm_sweeps=1, record_trace=True,
                 max_sweeps=nsw + 1)
    c.chain_run_sweeps(nsw + 1)
    assert not np.array_equal(c.chain_get_trace()["weight"], tr["weight"])
    c.close()


def test_chain_reports_nonconvergence():
    """A Lanczos run that hits its step cap must surface as FKMC_ERR_NOCONV from fkmc_chain_init / fkmc_chain_run_sweeps (not be
    consumed silently as a Metropolis weight), and the flag must not leak into later calls."""
    c = fk.Context("cubic2d", 16, max_batch=8)
    c.set_option("lanczos_max_steps", 16)
    with pytest.raises(fk.FkmcError) as ei:
        c.chain_init(8, 10.0, 2.0, cheb_moves=True, max_sweeps=4)
    assert ei.value.code == 4
    c.set_option("lanczos_max_steps", 0)
    c.chain_init(8, 10.0, 2.0, cheb_moves=True, max_sweeps=4)
    c.chain_run_sweeps(1)
    c.set_option("lanczos_max_steps", 16)
    with pytest.raises(fk.FkmcError) as ei:
        c.chain_run_sweeps(1)
    assert ei.value.code == 4
    c.set_option("lanczos_max_steps", 0)
    c.chain_run_sweeps(1)                                   # flag was cleared by the failing call
    f = c.chain_get_state()["f"]
    c.logz_kpm(f, 2.0, 1.0, 10.0, 12, 24)                   # and does not surface in an unrelated evaluator call
    c.close()
    c = fk.Context("cubic3d", 4, max_batch=2)               # the one-kernel KPM path (csrc/kpm.cu) has the same cap
    c.set_option("lanczos_max_steps", 8)
    with pytest.raises(fk.FkmcError) as ei:
        c.chain_init(2, 5.0, 4.0, cheb_moves=True, max_sweeps=2)
    assert ei.value.code == 4
    c.close()


# ---------------- per-sweep histories (measure_spectrum, measure_spectrum_history, measure_focc, measure_ipr) ----------------
@pytest.mark.parametrize("cheb", [False, True])
def test_chain_histories_match_oracle(cheb):
    nch, nsw, L, U, beta = 3, 4, 8, 4.0, 4.0
    c = fk.Context("cubic2d", L, max_batch=nch)
    c.chain_init(nch, beta, U, mc_flip=0.3, cheb_moves=cheb, seed=32167, sweep_len=16, ntherm_sweeps=1, max_sweeps=nsw + 1,
                 measure_history=True, measure_ipr=True)
    c.chain_run_sweeps(2)
    c.chain_run_sweeps(nsw - 1)
    h, se = c.chain_get_history(), c.chain_get_series()
    assert h["n_measured"] == nsw and se["n_measured"] == nsw
    for ch in range(nch):
        p = o.make_params(kind=o.CUBIC2D, L=L, beta=beta, U=U, mc_flip=0.3, cheb_moves=cheb, seed=32167, nsweeps=nsw, sweep_len=16,
                          ntherm_sweeps=1, measure_ipr=True)
        r = o.mc_run(p, rank=ch, trace=False)
        sh, fo = o.mc_histories(p, rank=ch)
        scale = np.abs(sh).max()
        assert np.array_equal(h["focc_history"][:, ch], fo)                                      # src/measures/focc_history.cpp:7-12
        assert np.abs(h["spectrum_history"][:, ch] - sh).max() <= TOL * scale                    # src/measures/spectrum_history.cpp:13-19
        assert np.abs(h["spectrum_mean"][ch] - r["spectrum_avg"]).max() <= TOL * scale           # src/measures/spectrum.cpp:13-21
        assert np.abs(se["energies"][:, ch] - r["energies"]).max() <= 1e-9 * np.abs(r["energies"]).max()
        for m in range(nsw):                                                                      # include/fk_mc/measures/ipr.hpp:39-56
            ev = sh[m]
            gaps = np.minimum(np.diff(ev, prepend=-np.inf), np.diff(ev, append=np.inf))
            iso = gaps > 1e-6 * scale  # the IPR is basis dependent inside (near-)degenerate subspaces
            assert iso.sum() > 32
            assert np.abs(h["ipr_history"][m, ch][iso] - r["ipr_history"][m][iso]).max() <= 1e-7
    c.close()


@pytest.mark.parametrize("cheb", [False, True])
def test_eigenfunctions_history(cheb):
    """measure_eigenfunctions (src/measures/eigenfunctions.cpp:12-18): the N x N eigenvector matrix of every chain at every measured sweep
    diagonalises the Hamiltonian of that sweep's configuration (focc_history) with the eigenvalues of spectrum_history."""
    nch, nsw, L, U, beta = 2, 3, 8, 4.0, 4.0
    c = fk.Context("cubic2d", L, max_batch=nch)
    c.chain_init(nch, beta, U, cheb_moves=cheb, seed=9, sweep_len=16, ntherm_sweeps=0, max_sweeps=nsw, measure_history=True,
                 measure_eigenfunctions=True, measure_ipr=True)
    c.chain_run_sweeps(nsw)
    V, h = c.chain_get_eigenfunctions(), c.chain_get_history()
    assert V.shape == (nsw, nch, L * L, L * L)
    H0 = o.hopping_dense(o.CUBIC2D, L)
    for m in range(nsw):
        for ch in range(nch):
            H = H0 + np.diag(U * h["focc_history"][m, ch] - U / 2)
            ev, W = h["spectrum_history"][m, ch], V[m, ch]
            assert np.abs(H @ W - W * ev).max() <= 1e-10 * np.abs(ev).max()
            assert np.abs(W.T @ W - np.eye(L * L)).max() <= 1e-9
            assert np.allclose(h["ipr_history"][m, ch], (np.sum(W ** 4, axis=0) ** 0.25) / np.sum(W ** 2, axis=0), rtol=1e-10)
    c.close()


def test_fsector_series():
    """nf0 / nfpi series (measure_nf0pi): recomputed from the focc history with the reference's staggered phase (-1)^(x+y)."""
    nch, nsw, L = 3, 4, 8
    c = fk.Context("cubic2d", L, max_batch=nch)
    c.chain_init(nch, 4.0, 4.0, mc_flip=0.3, seed=5, sweep_len=16, ntherm_sweeps=1, max_sweeps=nsw + 1, measure_history=True)
    c.chain_run_sweeps(nsw + 1)
    fs, h = c.chain_get_fsector(), c.chain_get_history()
    idx = np.arange(L * L)
    phase = np.where(((idx // L) + (idx % L)) % 2 == 0, 1, -1)
    assert fs["n_measured"] == nsw
    assert np.array_equal(fs["nf0"], h["focc_history"].sum(axis=2))
    assert np.array_equal(fs["nfpi"], np.abs((h["focc_history"] * phase).sum(axis=2)))
    assert np.array_equal(fs["nf0"], c.chain_get_series()["nf"])
    c.close()
    c3 = fk.Context("cubic3d", 4, max_batch=2)
    c3.chain_init(2, 2.0, 2.0, seed=5, sweep_len=16, ntherm_sweeps=0, max_sweeps=2, measure_history=True)
    c3.chain_run_sweeps(2)
    i3 = np.arange(64)
    ph3 = np.where(((i3 // 16) + (i3 // 4) % 4 + i3 % 4) % 2 == 0, 1, -1)
    assert np.array_equal(c3.chain_get_fsector()["nfpi"], np.abs((c3.chain_get_history()["focc_history"] * ph3).sum(axis=2)))
    c3.close()


def test_histories_off_is_a_state_error():
    c = fk.Context("cubic2d", 8, max_batch=1)
    c.chain_init(1, 1.0, 1.0, max_sweeps=2)
    c.chain_run_sweeps(2)
    h = c.chain_get_history()
    assert h["spectrum_history"] is None and h["ipr_history"] is None and h["spectrum_mean"].shape == (1, 64)
    import ctypes as C
    buf = np.zeros((2, 1, 64))
    rc = c.lib.fkmc_chain_get_history(c.h, None, None, buf.ctypes.data_as(C.POINTER(C.c_double)), None, None)
    assert rc == 5  # FKMC_ERR_STATE
    c.close()


def test_binned_observables_within_jackknife_error():
    """north_star: binned observables (energy, cv, IPR) agree with the reference's within its jackknife error.  GPU series of 16 chains
    x 48 sweeps and the oracle's for the same ranks go through the same analysis (binning of E, jackknife of cv: prog/data_save.hxx:158-199)."""
    from fk_mc_b200 import stats
    nch, nsw, L, U, beta = 16, 48, 8, 4.0, 2.0
    c = fk.Context("cubic2d", L, max_batch=nch)
    c.chain_init(nch, beta, U, seed=777, sweep_len=16, ntherm_sweeps=4, max_sweeps=nsw + 4, measure_history=True, measure_ipr=True)
    c.chain_run_sweeps(nsw + 4)
    se, h = c.chain_get_series(), c.chain_get_history()
    e_ref, d2_ref, ipr_ref = np.zeros((nsw, nch)), np.zeros((nsw, nch)), np.zeros((nsw, nch))
    for ch in range(nch):
        p = o.make_params(kind=o.CUBIC2D, L=L, beta=beta, U=U, seed=777, nsweeps=nsw, sweep_len=16, ntherm_sweeps=4, measure_ipr=True)
        r = o.mc_run(p, rank=ch, trace=False)
        e_ref[:, ch], d2_ref[:, ch] = r["energies"], r["d2energies"]
        ipr_ref[:, ch] = r["ipr_history"].mean(axis=1)
    ipr_gpu = h["ipr_history"].mean(axis=2)
    depth = 4
    rg = stats.energy_report(se["energies"], se["d2energies"], beta, L * L, max_depth=depth)
    rr = stats.energy_report(e_ref, d2_ref, beta, L * L, max_depth=depth)
    for name in ("energy", "d2energy", "cv"):
        (_, mg, _, eg), (_, mr, _, er) = rg[name]["stats"], rr[name]["stats"]
        assert er > 0 and abs(mg - mr) <= er, name            # within the reference's jackknife / binning error
        assert abs(eg - er) <= 0.05 * er, name
    ig = stats.accumulate_binning(stats.pool_chains(ipr_gpu)[::-1], depth)
    ir = stats.accumulate_binning(stats.pool_chains(ipr_ref)[::-1], depth)
    b = stats.estimate_bin(ir)
    assert ir[b][3] > 0 and abs(ig[b][1] - ir[b][1]) <= ir[b][3]
    c.close()


# ---------------- fast update of the dense moves: rank-one secular solver + tracked eigenvectors ----------------
@pytest.mark.parametrize("n", [2, 7, 64, 100, 256, 576, 1024])
def test_secular_update_stage(ctx8, n):
    """eig(diag(lam) + rho z z^T) against LAPACK: random problems of both signs, tiny and exactly zero components, exactly
    degenerate and nearly degenerate poles, rho = 0."""
    rng = np.random.default_rng(n)
    B = 8
    lam = np.sort(rng.normal(size=(B, n)) * 2, axis=1)
    z = rng.normal(size=(B, n))
    z /= np.linalg.norm(z, axis=1, keepdims=True)
    rho = np.array([2.0, -2.0, 0.37, -8.0, 1.0, -1.0, 0.0, 4.0])
    if n > 4:
        z[2, ::3] = 0.0                        # exact zeros: those poles stay eigenvalues
        z[3, 1::2] *= 1e-12                    # tiny components
        lam[4, n // 2:n // 2 + 3] = lam[4, n // 2]   # exactly degenerate poles
        lam[5, 1:] = np.sort(lam[5, 0] + np.cumsum(np.abs(rng.normal(size=n - 1)) * 1e-9))  # a cluster with gaps ~1e-9
        z[2] /= np.linalg.norm(z[2]); z[3] /= np.linalg.norm(z[3])
    got = ctx8.secular_update(lam, z, rho)
    for b in range(B):
        ref = sl.eigvalsh(np.diag(lam[b]) + rho[b] * np.outer(z[b], z[b]))
        scale = max(1.0, np.abs(ref).max())
        assert np.abs(got[b] - ref).max() <= 1e-13 * scale * max(1, n // 64), (n, b)
        assert (np.diff(got[b]) >= 0).all()


FAST_CASES = [("cubic2d", 8, 1.0, 1.0, 0.0), ("cubic2d", 8, 4.0, 4.0, 0.5), ("cubic2d", 8, 4.0, 4.0, 1.0), ("cubic2d", 16, 2.0, 10.0, 0.3),
              ("cubic3d", 4, 4.0, 5.0, 0.5), ("triangular", 6, 2.0, 10.0, 0.0), ("honeycomb", 6, 2.0, 10.0, 0.3), ("cubic1d", 16, 2.0, 3.0, 0.5)]


@pytest.mark.parametrize("kind,L,U,beta,flip", FAST_CASES)
def test_fast_update_chain_matches_oracle(kind, L, U, beta, flip):
    """The secular fast-update path must give the oracle's (i.e. the full eigensolve's) weights, accept/reject sequence, energies and
    final spectrum; refresh every 2 sweeps exercises the re-diagonalisation and its consistency check."""
    nch, nsw, sl_ = 4, 5, 16
    c = fk.Context(kind, L, max_batch=nch)
    add = 0.0 if flip == 1.0 else 1.0
    c.chain_init(nch, beta, U, mc_flip=flip, mc_add_remove=add, seed=32167, sweep_len=sl_, ntherm_sweeps=1, measure_energy=True,
                 record_trace=True, max_sweeps=nsw + 1, fast_update=True, fu_refresh_sweeps=2)
    c.chain_run_sweeps(nsw + 1)
    tr, se, st = c.chain_get_trace(), c.chain_get_series(), c.chain_get_state(spectrum=True)
    for ch in range(nch):
        p = o.make_params(kind=o.KINDS[kind], L=L, beta=beta, U=U, mc_flip=flip, mc_add_remove=add, seed=32167, nsweeps=nsw, sweep_len=sl_,
                          ntherm_sweeps=1)
        r = o.mc_run(p, rank=ch)
        t = r["trace"]
        wtol = max(1e-9, TOL * float(np.abs(t["logz_new"]).max()))
        assert np.array_equal(t["u"], tr["u"][:, ch]) and np.array_equal(t["site_a"], tr["site_a"][:, ch])
        assert (np.abs(t["weight"] - tr["weight"][:, ch]) <= wtol * np.maximum(1.0, np.abs(t["weight"]))).all()
        assert np.array_equal(t["accepted"], tr["accepted"][:, ch])
        assert np.array_equal(r["f_final"], st["f"][ch]) and r["naccept"] == st["naccept"][ch]
        assert np.abs(r["energies"] - se["energies"][:, ch]).max() <= 1e-9 * max(1.0, np.abs(r["energies"]).max())
        assert np.abs(r["d2energies"] - se["d2energies"][:, ch]).max() <= 1e-9 * max(1.0, np.abs(r["d2energies"]).max())
        ref = o.calc_ed(o.KINDS[kind], L, r["f_final"], U, U / 2, beta)["spectrum"]
        assert np.abs(st["spectrum"][ch] - ref).max() <= TOL * np.abs(ref).max()
    c.close()


def test_fast_update_equals_full_solve_at_baseline_size():
    """c2 shape (cubic2d L=16, beta=10, U=2): 64 chains x 12 sweeps without any refresh -- the tracked spectra stay within 1e-11 of a
    fresh eigensolve of the final configurations, and the accept sequence equals the full-solve path's."""
    nch, nsw, L, U, beta = 64, 12, 16, 2.0, 10.0
    res = []
    for fast in (False, True):
        c = fk.Context("cubic2d", L, max_batch=nch)
        c.chain_init(nch, beta, U, mc_flip=0.2, seed=4242, sweep_len=16, ntherm_sweeps=0, record_trace=True, max_sweeps=nsw, fast_update=fast,
                     fu_refresh_sweeps=1000)
        c.chain_run_sweeps(nsw)
        res.append((c.chain_get_trace(), c.chain_get_state(spectrum=True), c.chain_get_series()))
        if fast:
            fresh = c.logz_ed(res[-1][1]["f"], U, U / 2, beta)["spectrum"]
            assert np.abs(res[-1][1]["spectrum"] - fresh).max() <= 1e-11 * np.abs(fresh).max()
        c.close()
    (tf, sf, ef), (tq, sq, eq) = res
    assert np.array_equal(tf["accepted"], tq["accepted"]) and np.array_equal(sf["f"], sq["f"])
    assert np.abs(tf["weight"] - tq["weight"]).max() <= 1e-8 * max(1.0, np.abs(tf["weight"]).max())
    assert np.abs(ef["energies"] - eq["energies"]).max() <= 1e-9 * np.abs(ef["energies"]).max()
    assert tq["accepted"].mean() > 0.02  # some moves were accepted, i.e. the eigenvector update ran


def test_fast_update_ipr_history_from_tracked_eigenvectors():
    """With fast_update the per-sweep IPR comes from the tracked eigenvectors (no eigensolve): same values as the oracle's calc_ed(true)."""
    nch, nsw, L, U, beta = 3, 4, 8, 4.0, 4.0
    c = fk.Context("cubic2d", L, max_batch=nch)
    c.chain_init(nch, beta, U, mc_flip=0.3, seed=32167, sweep_len=16, ntherm_sweeps=1, max_sweeps=nsw + 1, measure_history=True, measure_ipr=True,
                 fast_update=True, fu_refresh_sweeps=1000)
    c.chain_run_sweeps(nsw + 1)
    h = c.chain_get_history()
    for ch in range(nch):
        p = o.make_params(kind=o.CUBIC2D, L=L, beta=beta, U=U, mc_flip=0.3, seed=32167, nsweeps=nsw, sweep_len=16, ntherm_sweeps=1, measure_ipr=True)
        r = o.mc_run(p, rank=ch, trace=False)
        sh, _ = o.mc_histories(p, rank=ch)
        assert np.abs(h["spectrum_history"][:, ch] - sh).max() <= TOL * np.abs(sh).max()
        for m in range(nsw):
            gaps = np.minimum(np.diff(sh[m], prepend=-np.inf), np.diff(sh[m], append=np.inf))
            iso = gaps > 1e-6 * np.abs(sh).max()
            assert iso.sum() > 32
            assert np.abs(h["ipr_history"][m, ch][iso] - r["ipr_history"][m][iso]).max() <= 1e-7
    c.close()


def test_fast_update_rejects_unsupported_setups():
    c = fk.Context("cubic2d", 8, max_batch=2)
    with pytest.raises(fk.FkmcError):
        c.chain_init(2, 1.0, 1.0, cheb_moves=True, fast_update=True)
    with pytest.raises(fk.FkmcError):
        c.chain_init(2, 1.0, 1.0, mc_reshuffle=0.1, fast_update=True)
    c.close()


def test_chain_golden_trace(golden):
    for t in golden["mc_trace"]["traces"]:
        c = fk.Context("cubic2d", 8, max_batch=2)
        c.chain_init(2, t["beta"], t["U"], mc_flip=t["mc_flip"], mc_reshuffle=t["mc_reshuffle"], cheb_moves=t["cheb"], seed=32167,
                     sweep_len=16, ntherm_sweeps=1, record_trace=True, max_sweeps=4)
        c.chain_run_sweeps(4)
        tr, se, st = c.chain_get_trace(), c.chain_get_series(), c.chain_get_state()
        ch = t["rank"]
        assert tr["accepted"][:, ch].tolist() == t["accepted"] and tr["site_a"][:, ch].tolist() == t["site_a"]
        assert np.array_equal(tr["u"][:, ch], np.array(t["u"]))
        assert np.allclose(tr["weight"][:, ch], t["weight"], rtol=1e-9, atol=1e-12)
        assert np.allclose(se["energies"][:, ch], t["energies"], rtol=1e-10)
        assert st["f"][ch].tolist() == t["f_final"]
        c.close()


def test_chain_offset_is_rank_seed():
    # chain c of a context with chain0 = k is the reference's MPI rank k + c (mc_metropolis.cpp:25)
    a = fk.Context("cubic2d", 8, max_batch=4)
    b = fk.Context("cubic2d", 8, max_batch=2)
    a.chain_init(4, 4.0, 4.0, seed=100, max_sweeps=2)
    b.chain_init(2, 4.0, 4.0, seed=100, chain0=2, max_sweeps=2)
    a.chain_run_sweeps(2)
    b.chain_run_sweeps(2)
    sa, sb = a.chain_get_state(), b.chain_get_state()
    assert np.array_equal(sa["f"][2:], sb["f"]) and np.array_equal(sa["logZ"][2:], sb["logZ"])
    assert np.array_equal(a.chain_get_series()["energies"][:, 2:], b.chain_get_series()["energies"])
    a.close()
    b.close()


def test_chain_flip_early_out_on_empty_config():
    # move_flip returns 0 without drawing when the configuration is empty (moves.cpp:8): only u is drawn each step
    c = fk.Context("cubic2d", 8, max_batch=1)
    c.chain_init(1, 1.0, 1.0, mc_flip=1.0, mc_add_remove=0.0, nf_start=0, seed=5, sweep_len=8, record_trace=True, max_sweeps=1)
    c.chain_run_sweeps(1)
    tr = c.chain_get_trace()
    # nf_start = 0 makes randomize_f draw nf itself first (configuration.cpp:49); compare with the oracle end to end
    p = o.make_params(kind=o.CUBIC2D, L=8, beta=1.0, U=1.0, mc_flip=1.0, mc_add_remove=0.0, nf_start=0, seed=5, nsweeps=0, sweep_len=8,
                      ntherm_sweeps=1)
    r = o.mc_run(p)
    assert np.array_equal(r["trace"]["u"], tr["u"][:, 0]) and np.array_equal(r["trace"]["accepted"], tr["accepted"][:, 0])
    c.close()


def test_chain_full_batch_statistics():
    # 1024 independent chains (BASELINE config 1 shape): mean energy agrees with the oracle's chains within 5 sigma
    nch, nsw = 1024, 6
    c = fk.Context("cubic2d", 8, max_batch=nch)
    c.chain_init(nch, 1.0, 1.0, seed=32167, sweep_len=16, ntherm_sweeps=2, max_sweeps=nsw + 2)
    c.chain_run_sweeps(nsw + 2)
    se = c.chain_get_series()
    assert se["n_measured"] == nsw and np.isfinite(se["energies"]).all()
    ref = []
    for ch in range(16):
        p = o.make_params(kind=o.CUBIC2D, L=8, beta=1.0, U=1.0, seed=32167, nsweeps=nsw, sweep_len=16, ntherm_sweeps=2)
        r = o.mc_run(p, rank=ch, trace=False)
        assert np.abs(r["energies"] - se["energies"][:, ch]).max() <= 1e-9 * np.abs(r["energies"]).max()
        ref.append(r["energies"].mean())
    m, s = se["energies"].mean(), se["energies"].mean(axis=0).std() / math.sqrt(nch)
    assert abs(m - np.mean(ref)) <= 5 * (s + np.std(ref) / 4)
    assert c.launch_count() > 0
    c.close()


# ---------------- C++ host mirror of the reference API (include/fk_mc_b200/fk_mc.hpp) ----------------
@pytest.mark.parametrize("cheb,flip", [(0, 0.0), (0, 0.5), (1, 0.5)])
def test_cpp_host_api_matches_oracle(tmp_path, cheb, flip):
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "host_api_test")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(root, "include"), os.path.join(root, "tests", "cpp", "host_api_test.cpp"),
                           "-o", exe, "-L" + os.path.join(root, "fk_mc_b200", "lib"), "-lfkmc_b200",
                           "-Wl,-rpath," + os.path.join(root, "fk_mc_b200", "lib")])
    L, beta, U, nsw, seed, rank = 8, 4.0, 4.0, 3, 32167, 2
    out = subprocess.check_output([exe, str(L), str(beta), str(U), str(cheb), str(flip), str(nsw), str(seed), str(rank)]).decode()
    rows = {ln.split()[0]: ln.split()[1:] for ln in out.strip().splitlines()}
    assert "exception" not in rows, out
    # fk_mc<hypercubic_lattice<2>>: define_parameters -> initialize -> run (one chain, batch of one per evaluation) ...
    p = o.make_params(kind=o.CUBIC2D, L=L, beta=beta, U=U, mc_flip=flip, cheb_moves=bool(cheb), seed=seed, nsweeps=nsw, sweep_len=16,
                      ntherm_sweeps=1)
    r = o.mc_run(p, rank=rank)
    assert int(rows["naccept"][0]) == r["naccept"]
    assert np.array_equal(np.array(rows["f"], dtype=np.int32), r["f_final"])
    assert np.allclose(np.array(rows["energies"], dtype=float), r["energies"], rtol=1e-10)
    assert np.allclose(np.array(rows["d2energies"], dtype=float), r["d2energies"], rtol=1e-9)
    assert np.allclose(np.array(rows["spectrum_mean"], dtype=float), r["spectrum_avg"], rtol=0, atol=1e-10 * np.abs(r["spectrum_avg"]).max())
    assert rows["history_shape"] == [str(L * L), str(nsw), str(L * L)]      # [index][measurement] like observables_t
    # ... and fk_mc::run_batched: ranks `rank` and `rank + 1` as two device-resident chains
    for c in range(2):
        rb = o.mc_run(p, rank=rank + c)
        assert int(rows["b%d_naccept" % c][0]) == rb["naccept"]
        assert np.array_equal(np.array(rows["b%d_f" % c], dtype=np.int32), rb["f_final"])
        assert np.allclose(np.array(rows["b%d_energies" % c], dtype=float), rb["energies"], rtol=1e-10)
        assert np.allclose(np.array(rows["b%d_spectrum_mean" % c], dtype=float), rb["spectrum_avg"], rtol=0, atol=1e-10 * np.abs(rb["spectrum_avg"]).max())
    assert rows["b_history_shape"] == [str(L * L), str(nsw), str(L * L)]
    assert rows["b_stiffness_shape"] == [str(nsw), "2", str(nsw)] and len(rows["b0_stiffness"]) == nsw
    assert rows["mismatch_throws"] == ["1"] and rows["honeycomb_odd_throws"] == ["1"] and rows["no_moves_throws"] == ["1"]


# ---------------- eigenvector path: calc_ed(true), measure_ipr ----------------
@pytest.mark.parametrize("kind,L,U", [("cubic2d", 8, 4.0), ("cubic2d", 16, 2.0), ("triangular", 24, 2.0), ("cubic3d", 4, 0.37)])
def test_eigenvectors_and_ipr(kind, L, U):
    beta = 3.0
    c = fk.Context(kind, L, max_batch=3)
    n = c.N
    fs = np.stack([o.randomize_f(32167 + i, n, n // 2)[0] for i in range(3)])
    r = c.eigh(fs, U, U / 2, beta)
    ri = c.ipr(fs, U, U / 2, beta)
    for b in range(3):
        H = o.hopping_dense(o.KINDS[kind], L) + np.diag(U * fs[b] - U / 2)
        ev, V = r["spectrum"][b], r["evecs"][b]
        ref = o.calc_ed(o.KINDS[kind], L, fs[b], U, U / 2, beta, vectors=True)
        scale = np.abs(ref["spectrum"]).max()
        assert np.abs(ev - ref["spectrum"]).max() <= TOL * scale
        assert np.abs(H @ V - V * ev).max() <= 1e-10 * scale          # residual
        assert np.abs(V.T @ V - np.eye(n)).max() <= 1e-9               # orthonormal columns
        assert abs(r["logZ"][b] - ref["logZ"]) <= TOL * abs(ref["logZ"])
        # IPR (ipr.hpp:47-53) is basis independent for non-degenerate states: compare with the oracle's eigenvectors
        ipr_ref = o.measure_ipr(ref["evecs"])
        gaps = np.minimum(np.diff(ev, prepend=-np.inf), np.diff(ev, append=np.inf))
        iso = gaps > 1e-6 * scale
        assert iso.sum() > n // 2
        assert np.abs(ri["ipr"][b][iso] - ipr_ref[iso]).max() <= 1e-7
        assert np.allclose(ri["ipr"][b], (np.sum(V ** 4, axis=0) ** 0.25) / np.sum(V ** 2, axis=0), rtol=1e-10)
    c.close()


def test_eigenvectors_degenerate_spectrum():
    # free lattice: massively degenerate; any orthonormal basis of each eigenspace is acceptable
    c = fk.Context("cubic2d", 8)
    r = c.eigh(np.zeros(64, np.int32), 1.0, 0.3, 2.0)
    H = o.hopping_dense(o.CUBIC2D, 8) - 0.3 * np.eye(64)
    V, ev = r["evecs"][0], r["spectrum"][0]
    assert np.abs(H @ V - V * ev).max() <= 1e-9
    assert np.abs(V.T @ V - np.eye(64)).max() <= 1e-8
    c.close()


def test_chain_ipr_matches_oracle_history():
    c = fk.Context("cubic2d", 8, max_batch=2)
    c.chain_init(2, 4.0, 4.0, seed=32167, sweep_len=16, ntherm_sweeps=0, max_sweeps=2)
    c.chain_run_sweeps(2)
    got = c.chain_ipr()
    for ch in range(2):
        p = o.make_params(kind=o.CUBIC2D, L=8, beta=4.0, U=4.0, seed=32167, nsweeps=2, sweep_len=16, ntherm_sweeps=0, measure_ipr=True)
        r = o.mc_run(p, rank=ch, trace=False)
        ev = got["spectrum"][ch]
        gaps = np.minimum(np.diff(ev, prepend=-np.inf), np.diff(ev, append=np.inf))
        iso = gaps > 1e-6 * np.abs(ev).max()  # the IPR is basis dependent inside (near-)degenerate subspaces
        assert iso.sum() > 32
        assert np.abs(got["ipr"][ch][iso] - r["ipr_history"][-1][iso]).max() <= 1e-7
    c.close()


# ---------------- stiffness: the reference's only golden values that involve eigenvectors (test/stiffness_test.cpp:60-63) ----------------
F5 = [0, 1, 0, 1, 1, 0, 1, 0, 0, 1, 0, 0, 1, 0, 1, 0, 1, 1, 1, 0, 1, 1, 0, 1, 1]
F7 = [0, 1, 0, 0, 0, 0, 1, 1, 0, 1, 1, 0, 0, 1, 1, 0, 1, 1, 0, 1, 1, 0, 1, 0, 0, 0, 1, 1, 0, 1, 1, 1, 1, 1, 0, 1, 1, 0, 1, 1, 1, 1, 0, 0, 0, 1, 0, 1, 1]


@pytest.mark.parametrize("U,f,golden", [(0.0, F5, 1.31597), (2.0, F5, 0.287547), (6.0, F5, 0.00260831), (0.37, F7, 1.26046)])
def test_stiffness_reference_goldens(U, f, golden):
    from fk_mc_b200 import measures
    L = int(round(len(f) ** 0.5))
    c = fk.Context("cubic2d", L)
    st, cond = measures.stiffness(c, np.array(f, np.int32), U, U / 2, 1000.0)
    assert st[0] == pytest.approx(golden, abs=1e-3)           # the reference's tolerance
    ref, cref = o.stiffness(o.CUBIC2D, L, np.array(f, np.int32), U, U / 2, 1000.0)
    assert st[0] == pytest.approx(ref, abs=1e-8)
    assert cond[0, 0] == pytest.approx(cref[0], rel=1e-6, abs=1e-9)
    c.close()


@pytest.mark.parametrize("kind,L,ndim", [("cubic2d", 8, 2), ("cubic3d", 4, 3), ("cubic2d", 16, 2)])
def test_stiffness_gpu_contraction_matches_host(kind, L, ndim):
    """fkmc_stiffness_batched (V^T Jm V on DMMA, Kubo sums on the device) against the numpy contraction of the same GPU eigenvectors and
    against the oracle, ten frequencies (more than one accumulation pass), a batch of three configurations."""
    from fk_mc_b200 import measures
    U, beta = 2.0, 5.0
    c = fk.Context(kind, L, max_batch=3)
    n = c.N
    fs = np.stack([o.randomize_f(7 + i, n, n // 2)[0] for i in range(3)])
    wg = np.linspace(-3.0, 3.0, 10)
    st, cd = c.stiffness(fs, U, U / 2, beta, offset=0.05, wgrid=wg)
    sh, ch = measures.stiffness_host_contraction(c, fs, U, U / 2, beta, ndim=ndim, offset=0.05, wgrid=wg)
    assert np.abs(st - sh).max() <= 1e-9 * max(1.0, np.abs(sh).max())
    assert np.abs(cd - ch).max() <= 1e-8 * max(1.0, np.abs(ch).max())
    if ndim == 2:
        for b in range(3):
            ref, cref = o.stiffness(o.KINDS[kind], L, fs[b], U, U / 2, beta, offset=0.05, wgrid=wg)
            assert st[b] == pytest.approx(ref, abs=1e-8)
            assert np.abs(cd[b] - cref).max() <= 1e-6 * max(1.0, np.abs(cref).max())
    c.close()


@pytest.mark.parametrize("cheb", [False, True])
def test_chain_measure_stiffness_series(cheb):
    """Chain parameter measure_stiffness (fk_mc.hxx:101-105, stiffness.hpp:129-187): the stiffness / conductivity of every measured
    sweep's configuration (from focc_history) equals the oracle's measure on that configuration."""
    L, U, beta, n_chains, nsw = 6, 2.0, 4.0, 3, 4
    wg = [0.0, 0.5, 1.0]
    c = fk.Context("cubic2d", L, max_batch=n_chains)
    c.chain_init(n_chains, beta, U, U / 2, U / 2, seed=77, sweep_len=5, max_sweeps=nsw, cheb_moves=cheb, measure_history=True,
                 measure_stiffness=True, cond_wgrid=wg, cond_offset=0.07)
    c.chain_run_sweeps(nsw)
    r = c.chain_get_stiffness()
    h = c.chain_get_history()
    nm = r["n_measured"]
    assert nm == h["n_measured"] >= 3 and r["stiffness"].shape == (nm, n_chains) and r["cond"].shape == (nm, n_chains, 3)
    for m in range(nm):
        for k in range(n_chains):
            f = h["focc_history"][m, k].astype(np.int32)
            ref, cref = o.stiffness(o.CUBIC2D, L, f, U, U / 2, beta, offset=0.07, wgrid=np.array(wg))
            assert r["stiffness"][m, k] == pytest.approx(ref, abs=1e-8)
            assert np.abs(r["cond"][m, k] - cref).max() <= 1e-6 * max(1.0, np.abs(cref).max())
    c.close()
    c = fk.Context("triangular", 6)
    with pytest.raises(fk.FkmcError):
        c.chain_init(1, beta, U, U / 2, U / 2, measure_stiffness=True)
    c.close()


def test_stiffness_rejects_other_lattices():
    c = fk.Context("triangular", 6)
    with pytest.raises(fk.FkmcError):
        c.stiffness(np.zeros(36, np.int32), 1.0, 0.5, 1.0)
    c.close()


# ---------------- local KPM re-evaluation (kpm2d.cu) ----------------
@pytest.mark.parametrize("kind,L,M", [("cubic2d", 32, 16), ("cubic2d", 16, 12), ("triangular", 24, 14)])
def test_kpm_local_matches_full(kind, L, M):
    """fkmc_logz_kpm_batched_local: configurations that differ from a reference in one site (add / remove) or two (flip) re-evaluated from
    the reference's trace-sum record, against the full evaluation and the oracle; a chain of 40 accepted single-site changes without a
    full evaluation in between; fall-backs: many changed sites, a record of another M, no reference."""
    U, beta, B = 2.0, 8.0, 6
    G = 2 * M
    c = fk.Context(kind, L, max_batch=B)
    n = c.N
    rng = np.random.default_rng(17)
    f0 = np.stack([o.randomize_f(50 + i, n, n // 2)[0] for i in range(B)])
    r0 = c.logz_kpm_local(f0, U, U / 2, beta, M, G)                        # no reference: full, returns the records
    full0 = c.logz_kpm(f0, U, U / 2, beta, M, G)
    assert np.array_equal(r0["logZ"], full0["logZ"]) and np.all(r0["state"][:, 54] == 1.0)
    f1 = f0.copy()
    for b in range(B):
        if b % 2 == 0:
            f1[b, rng.integers(n)] ^= 1                                       # add / remove
        else:
            occ, emp = np.flatnonzero(f1[b] == 1), np.flatnonzero(f1[b] == 0)  # flip: move one f-electron
            f1[b, rng.choice(occ)] = 0
            f1[b, rng.choice(emp)] = 1
    r1 = c.logz_kpm_local(f1, U, U / 2, beta, M, G, f_ref=f0, state_ref=r0["state"])
    full1 = c.logz_kpm(f1, U, U / 2, beta, M, G)
    assert np.abs(r1["logZ"] - full1["logZ"]).max() <= 1e-12 * np.abs(full1["logZ"]).max()
    assert np.abs(r1["moments"] - full1["moments"]).max() <= 1e-12
    assert np.array_equal(r1["a"], full1["a"]) and np.array_equal(r1["b"], full1["b"])
    ref = o.calc_chebyshev(o.KINDS[kind], L, f1[0], U, U / 2, beta, M, G)
    assert abs(r1["logZ"][0] - ref["logZ"]) <= TOL * abs(ref["logZ"])
    # the record of a locally evaluated configuration serves as the next reference: 40 steps, then compare with a fresh full evaluation
    f, st = f1.copy(), r1["state"]
    for _ in range(40):
        g = f.copy()
        g[np.arange(B), rng.integers(n, size=B)] ^= 1
        r = c.logz_kpm_local(g, U, U / 2, beta, M, G, f_ref=f, state_ref=st)
        f, st = g, r["state"]
    fullN = c.logz_kpm(f, U, U / 2, beta, M, G)
    assert np.abs(r["logZ"] - fullN["logZ"]).max() <= 1e-11 * np.abs(fullN["logZ"]).max()
    assert np.abs(r["moments"] - fullN["moments"]).max() <= 1e-11
    # fall-backs are full evaluations: bit-identical to fkmc_logz_kpm_batched
    g = f.copy()
    g[:, :5] ^= 1
    rf = c.logz_kpm_local(g, U, U / 2, beta, M, G, f_ref=f, state_ref=st)
    assert np.array_equal(rf["logZ"], c.logz_kpm(g, U, U / 2, beta, M, G)["logZ"])
    g = f.copy()
    g[:, 7] ^= 1
    rm = c.logz_kpm_local(g, U, U / 2, beta, M - 2, 2 * (M - 2), f_ref=f, state_ref=st)   # the record was made for another M
    assert np.array_equal(rm["logZ"], c.logz_kpm(g, U, U / 2, beta, M - 2, 2 * (M - 2))["logZ"])
    c.close()


def test_kpm_local_chain_equals_full_chain():
    """Chebyshev-move chains with the local scheme on (default) and off: identical accept / reject sequences, logZ of every proposal within
    1e-12, also across a re-base (kpm_rebase_sweeps = 2) and with reshuffle moves (always full) mixed in."""
    out = []
    for loc in (1, 0):
        c = fk.Context("cubic2d", 16, max_batch=8)
        c.set_option("kpm_local", loc)
        c.set_option("kpm_rebase_sweeps", 2)
        c.chain_init(8, 6.0, 2.0, cheb_moves=True, seed=3, sweep_len=8, ntherm_sweeps=0, measure_energy=False, mc_flip=0.3, mc_add_remove=0.6,
                     mc_reshuffle=0.1, record_trace=True, max_sweeps=6)
        c.chain_run_sweeps(6)
        out.append((c.chain_get_trace(), c.chain_get_state()))
        c.close()
    (t1, s1), (t0, s0) = out
    assert np.array_equal(t1["accepted"], t0["accepted"]) and np.array_equal(s1["f"], s0["f"])
    assert np.abs(t1["logz_new"] - t0["logz_new"]).max() <= 1e-12 * np.abs(t0["logz_new"]).max()


def test_kpm_local_other_lattices_return_invalid_records():
    c = fk.Context("cubic3d", 6, max_batch=2)
    f = np.stack([o.randomize_f(1 + i, c.N, c.N // 2)[0] for i in range(2)])
    r = c.logz_kpm_local(f, 1.0, 0.5, 2.0, 8, 16)
    assert np.all(r["state"] == 0.0)
    g = f.copy()
    g[:, 3] ^= 1
    r2 = c.logz_kpm_local(g, 1.0, 0.5, 2.0, 8, 16, f_ref=f, state_ref=r["state"])
    assert np.array_equal(r2["logZ"], c.logz_kpm(g, 1.0, 0.5, 2.0, 8, 16)["logZ"])
    c.close()


def test_seam_calls_with_caller_owned_page_locked_buffers():
    """The evaluator seam with page-locked, reused host buffers (what a host loop holds; bench.py's e2e path): the reference configurations
    and records then travel on the library's second stream under the Lanczos kernel.  Results must be those of the plain calls, placed
    in the caller's arrays; a buffer of the wrong shape is refused."""
    import torch

    def pinned(*shape, dtype=torch.float64):
        return torch.zeros(shape, dtype=dtype).pin_memory().numpy()

    B, L, M, G = 6, 16, 12, 24
    c = fk.Context("cubic2d", L, max_batch=B)
    n = c.N
    f = pinned(B, n, dtype=torch.int32)
    f[:] = np.stack([o.randomize_f(3 + i, n, n // 2)[0] for i in range(B)])
    g = pinned(B, n, dtype=torch.int32)
    out = dict(moments=pinned(B, M), ab=pinned(B, 4), logZ=pinned(B), state=pinned(B, 64))
    ks = pinned(B, 64)
    r0 = c.logz_kpm_local(f, 2.0, 1.0, 5.0, M, G)
    ks[:] = r0["state"]
    for step in range(3):                      # the same buffers call after call
        g[:] = f
        g[np.arange(B), (7 * np.arange(B) + 11 * step) % n] ^= 1
        plain = c.logz_kpm_local(np.array(g), 2.0, 1.0, 5.0, M, G, f_ref=np.array(f), state_ref=np.array(ks))
        r = c.logz_kpm_local(g, 2.0, 1.0, 5.0, M, G, f_ref=f, state_ref=ks, out=out)
        assert r["logZ"] is out["logZ"] and r["state"] is out["state"] and r["moments"] is out["moments"]
        for k in ("logZ", "moments", "state", "e_min", "e_max"):
            assert np.array_equal(r[k], plain[k]), k
        full = c.logz_kpm(np.array(g), 2.0, 1.0, 5.0, M, G)
        assert np.abs(r["logZ"] - full["logZ"]).max() <= 1e-12 * np.abs(full["logZ"]).max()
        f[:] = g
        ks[:] = r["state"]
    oe = dict(spectrum=pinned(B, n), logZ=pinned(B))
    re_ = c.logz_ed(f, 2.0, 1.0, 5.0, out=oe)
    assert re_["spectrum"] is oe["spectrum"] and np.array_equal(re_["spectrum"], c.logz_ed(np.array(f), 2.0, 1.0, 5.0)["spectrum"])
    with pytest.raises(fk.FkmcError):
        c.logz_ed(f, 2.0, 1.0, 5.0, out=dict(spectrum=np.zeros((B, n + 1))))
    with pytest.raises(fk.FkmcError):
        c.logz_kpm_local(f, 2.0, 1.0, 5.0, M, G, out=dict(state=np.zeros((B, 64), dtype=np.float32)))
    c.close()


# ---------------- band path: folded lattice ordering, band -> band (sb2sb.cu) -> tridiagonal ----------------
@pytest.mark.parametrize("kind,L,U", [("cubic2d", 16, 2.0), ("cubic2d", 24, 0.5), ("cubic2d", 26, 4.0), ("cubic2d", 32, 1.0), ("triangular", 24, 2.0),
                                      ("triangular", 31, 2.0), ("honeycomb", 24, 2.0), ("honeycomb", 32, 1.0)])
def test_band_path_spectra(kind, L, U):
    """calc_ed(false) through the band path (sites reordered so that the lattice matrix has half-bandwidth <= 64, block bulge chasing on
    DMMA down to half-bandwidth 8) against the oracle and against the dense reduction of the same configurations; N = 676 is not a
    multiple of the tile, triangular L = 31 has the maximal bandwidth (2L + 1 = 63), ordered / empty / full configurations included."""
    beta = 7.0
    c = fk.Context(kind, L, max_batch=5)
    n = c.N
    fs = np.stack([o.randomize_f(11 + i, n, n // 2)[0] for i in range(2)] + [np.zeros(n, np.int32), np.ones(n, np.int32),
                                                                              (np.arange(n) % 2).astype(np.int32)])
    c.profile_enable(True)
    c.profile_reset()
    rb = c.logz_ed(fs, U, U / 2, beta, want_caches=True)
    assert c.profile_get("sb2sb")[1] >= 1 and c.profile_get("sy2sb")[1] == 0      # the band kernels did run
    c.set_option("band_path", 0)
    c.profile_reset()
    rd = c.logz_ed(fs, U, U / 2, beta)
    assert c.profile_get("sb2sb")[1] == 0
    c.profile_enable(False)
    for b in range(len(fs)):
        ref = o.calc_ed(o.KINDS[kind], L, fs[b], U, U / 2, beta)
        scale = np.abs(ref["spectrum"]).max()
        assert np.abs(rb["spectrum"][b] - ref["spectrum"]).max() <= TOL * scale
        assert np.abs(rb["spectrum"][b] - rd["spectrum"][b]).max() <= 1e-12 * scale
        assert abs(rb["logZ"][b] - ref["logZ"]) <= TOL * abs(ref["logZ"])
        assert np.abs(rb["cached_fermi"][b] - ref["cached_fermi"]).max() <= 1e-9
    c.close()


@pytest.mark.parametrize("kind,L,t,tp", [("cubic2d", 16, 0.7, 1.0), ("triangular", 16, 1.0, 0.45), ("triangular", 24, 1.3, -0.6), ("honeycomb_ref_lower", 24, 1.0, 1.0)])
def test_band_path_hoppings(kind, L, t, tp):
    """Band path with hopping constants other than 1 (triangular: t' on the diagonal bonds, which sit at distance 2L + 1 in the folded
    ordering) and for the literal lower-triangle honeycomb of the reference against the oracle."""
    U, beta = 1.5, 6.0
    c = fk.Context(kind, L, t=t, tp=tp, max_batch=2)
    n = c.N
    fs = np.stack([o.randomize_f(21 + i, n, n // 3)[0] for i in range(2)])
    c.profile_enable(True)
    r = c.logz_ed(fs, U, U / 2, beta)
    assert c.profile_get("sb2sb")[1] >= 1
    for b in range(2):
        ref = o.calc_ed(o.KINDS[kind], L, fs[b], U, U / 2, beta, t=t, tp=tp)
        assert np.abs(r["spectrum"][b] - ref["spectrum"]).max() <= TOL * np.abs(ref["spectrum"]).max()
        assert abs(r["logZ"][b] - ref["logZ"]) <= TOL * abs(ref["logZ"])
    c.close()


def test_band_path_selection():
    """The band path needs half-bandwidth <= 64 after folding and N >= band_min: cubic3d (2 L^2 = 128), triangular L = 32 (65) and small
    lattices stay on the dense reduction; the option band_min moves the threshold."""
    for kind, L, expect in [("cubic3d", 8, False), ("triangular", 32, False), ("cubic2d", 12, False), ("cubic2d", 16, True), ("cubic1d", 300, True)]:
        c = fk.Context(kind, L, max_batch=1)
        f = o.randomize_f(3, c.N, c.N // 2)[0]
        c.profile_enable(True)
        r = c.logz_ed(f, 1.0, 0.5, 2.0)
        assert (c.profile_get("sb2sb")[1] > 0) == expect, (kind, L)
        ref = o.calc_ed(o.KINDS[kind], L, f, 1.0, 0.5, 2.0)
        assert np.abs(r["spectrum"][0] - ref["spectrum"]).max() <= TOL * np.abs(ref["spectrum"]).max()
        if kind == "cubic2d" and L == 12:
            c.set_option("band_min", 64)
            c.profile_reset()
            r = c.logz_ed(f, 1.0, 0.5, 2.0)
            assert c.profile_get("sb2sb")[1] > 0
            assert np.abs(r["spectrum"][0] - ref["spectrum"]).max() <= TOL * np.abs(ref["spectrum"]).max()
        c.close()


def test_band_path_chain_equals_dense_chain():
    """Dense moves at N = 256 with the band path on and off: identical accept / reject sequences and final configurations."""
    L, U, beta, n_chains = 16, 3.0, 4.0, 6
    out = []
    for bp in (1, 0):
        c = fk.Context("cubic2d", L, max_batch=n_chains)
        c.set_option("band_path", bp)
        c.chain_init(n_chains, beta, U, U / 2, U / 2, mc_flip=0.5, mc_add_remove=0.5, seed=5, sweep_len=4, max_sweeps=5)
        c.chain_run_sweeps(5)
        st = c.chain_get_state()
        se = c.chain_get_series()
        out.append((st, se))
        c.close()
    assert np.array_equal(out[0][0]["f"], out[1][0]["f"]) and np.array_equal(out[0][0]["naccept"], out[1][0]["naccept"])
    assert np.allclose(out[0][1]["energies"], out[1][1]["energies"], rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("kind,L", [("honeycomb", 34), ("cubic2d", 36)])
def test_sizes_above_1024(kind, L):
    """N > 1024 (e.g. the N ~ 1152 honeycomb extension of BASELINE config 4): the one-stage blocked tridiagonalisation + bisection with
    several eigenvalues per thread; spectra, logZ, a two-chain run and the eigenvector path against the oracle."""
    U, beta = 2.0, 10.0
    c = fk.Context(kind, L, max_batch=2)
    n = c.N
    assert n > 1024
    fs = np.stack([o.randomize_f(32167 + i, n, n // 2)[0] for i in range(2)])
    r = c.logz_ed(fs, U, U / 2, beta, want_caches=True)
    for b in range(2):
        ref = o.calc_ed(o.KINDS[kind], L, fs[b], U, U / 2, beta)
        assert np.abs(r["spectrum"][b] - ref["spectrum"]).max() <= TOL * np.abs(ref["spectrum"]).max()
        assert abs(r["logZ"][b] - ref["logZ"]) <= TOL * abs(ref["logZ"])
        assert np.abs(r["cached_fermi"][b] - ref["cached_fermi"]).max() <= 1e-9
    k = c.logz_kpm(fs, U, U / 2, beta, *fk.cheb_sizes(n))
    kref = o.calc_chebyshev(o.KINDS[kind], L, fs[0], U, U / 2, beta, *fk.cheb_sizes(n))
    assert abs(k["logZ"][0] - kref["logZ"]) <= TOL * abs(kref["logZ"])
    ri = c.ipr(fs[:1], U, U / 2, beta)
    assert np.abs(ri["spectrum"][0] - r["spectrum"][0]).max() <= TOL * np.abs(r["spectrum"]).max()
    assert (ri["ipr"][0] > 0).all() and (ri["ipr"][0] <= 1.0 + 1e-12).all()
    c.chain_init(2, beta, U, seed=32167, sweep_len=4, ntherm_sweeps=0, record_trace=True, max_sweeps=1)
    c.chain_run_sweeps(1)
    tr = c.chain_get_trace()
    for ch in range(2):
        p = o.make_params(kind=o.KINDS[kind], L=L, beta=beta, U=U, seed=32167, nsweeps=1, sweep_len=4, ntherm_sweeps=0)
        t = o.mc_run(p, rank=ch)["trace"]
        assert np.array_equal(t["accepted"], tr["accepted"][:, ch]) and np.array_equal(t["site_a"], tr["site_a"][:, ch])
        assert np.abs(t["weight"] - tr["weight"][:, ch]).max() <= max(1e-9, TOL * np.abs(t["logz_new"]).max()) * max(1.0, np.abs(t["weight"]).max())
    c.close()


@pytest.mark.parametrize("kind,L,cheb,fast,flip", [("cubic2d", 8, True, False, 0.0), ("cubic2d", 16, False, False, 0.3), ("cubic2d", 16, False, True, 0.3),
                                                   ("cubic3d", 4, True, False, 0.0), ("cubic2d", 32, True, False, 0.0)])
def test_step_graph_equals_eager_launches(kind, L, cheb, fast, flip):
    """fkmc_chain_run_sweeps replays one captured CUDA graph per Metropolis step (no trace, no profiling); the chains, series and the
    launch count must be exactly those of the kernel-by-kernel path."""
    nch, nsw, U, beta = 6, 3, 2.0, 5.0
    out = []
    for graph in (1, 0):
        c = fk.Context(kind, L, max_batch=nch)
        c.set_option("cuda_graph", graph)
        c.chain_init(nch, beta, U, mc_flip=flip, cheb_moves=cheb, seed=11, sweep_len=16, ntherm_sweeps=0, max_sweeps=nsw, fast_update=fast,
                     fu_refresh_sweeps=2)
        l0 = c.launch_count()
        c.chain_run_sweeps(1)
        c.chain_run_sweeps(nsw - 1)
        out.append((c.chain_get_state(), c.chain_get_series(), c.launch_count() - l0))
        c.close()
    (sg, eg, lg), (se, ee, le) = out
    assert np.array_equal(sg["f"], se["f"]) and np.array_equal(sg["naccept"], se["naccept"]) and np.array_equal(sg["logZ"], se["logZ"])
    assert np.array_equal(eg["energies"], ee["energies"]) and np.array_equal(eg["d2energies"], ee["d2energies"])
    assert lg == le and lg > 0
