"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py (cpu_baseline / --impl reference) and
__graft_entry__.smoke().  The product package fk_mc_b200 never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "liboracle.so")

CUBIC1D, CUBIC2D, CUBIC3D, TRIANGULAR, HONEYCOMB, HONEYCOMB_REF, HONEYCOMB_REF_LOWER = 1, 2, 3, 4, 5, 6, 7
KINDS = {"cubic1d": 1, "cubic2d": 2, "cubic3d": 3, "triangular": 4, "honeycomb": 5, "honeycomb_ref": 6,
         "honeycomb_ref_lower": 7}

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class McParams(C.Structure):
    _fields_ = [("kind", C.c_int), ("L", C.c_int), ("t", C.c_double), ("tp", C.c_double), ("beta", C.c_double),
                ("U", C.c_double), ("mu_c", C.c_double), ("mu_f", C.c_double), ("mc_flip", C.c_double),
                ("mc_add_remove", C.c_double), ("mc_reshuffle", C.c_double), ("cheb_moves", C.c_int),
                ("cheb_prefactor", C.c_double), ("emode", C.c_int), ("seed", C.c_long), ("nf_start", C.c_int),
                ("nsweeps", C.c_int), ("sweep_len", C.c_int), ("ntherm_sweeps", C.c_int), ("measure_energy", C.c_int),
                ("measure_ipr", C.c_int), ("n_W", C.c_int), ("W", C.c_double * 8)]


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(LIB_PATH)
            for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".hpp"))):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_last_error.restype = C.c_char_p
        _lib.orc_ff_energy.restype = C.c_double
        _lib.orc_logz_from_spectrum.restype = C.c_double
        _lib.orc_cheb_moment.restype = C.c_double
    return _lib


def _ck(rc):
    if rc != 0:
        raise RuntimeError("oracle: " + lib().orc_last_error().decode())


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def lattice_size(kind, L):
    return lib().orc_lattice_size(kind, L)


def hopping_dense(kind, L, t=1.0, tp=1.0):
    n = lattice_size(kind, L)
    H = np.zeros((n, n))
    _ck(lib().orc_hopping_dense(kind, L, C.c_double(t), C.c_double(tp), _p(H, C.c_double)))
    return H


def index_to_pos(kind, L, index):
    pos = np.zeros(3, dtype=np.int32)
    _ck(lib().orc_index_to_pos(kind, L, index, _p(pos, C.c_int)))
    return pos


def randomize_f(seed, V, nf):
    f = np.zeros(V, dtype=np.int32)
    nxt = C.c_int(0)
    _ck(lib().orc_randomize_f(C.c_long(seed), V, nf, _p(f, C.c_int), C.byref(nxt)))
    return f, nxt.value & 0xFFFFFFFF


def rng_stream(seed, mode, V, count):
    out = np.zeros(count)
    _ck(lib().orc_rng_stream(C.c_long(seed), mode, V, count, _p(out, C.c_double)))
    return out


def ff_energy(f, W):
    f = np.ascontiguousarray(f, dtype=np.int32)
    W = np.ascontiguousarray(W, dtype=np.float64)
    return lib().orc_ff_energy(len(f), _p(f, C.c_int), len(W), _p(W, C.c_double))


def eigh(A, vectors=False):
    """A: symmetric (n, n) array; the oracle reads the lower triangle (column-major == row-major of A.T)."""
    n = A.shape[0]
    a = np.asfortranarray(A, dtype=np.float64)
    ev = np.zeros(n)
    vec = np.zeros((n, n), order="F") if vectors else None
    rc = lib().orc_eigh(n, a.ctypes.data_as(C.POINTER(C.c_double)), _p(ev, C.c_double),
                        None if vec is None else vec.ctypes.data_as(C.POINTER(C.c_double)))
    if rc != 0:
        raise RuntimeError("oracle eigh: no convergence")
    return (ev, vec) if vectors else ev


def tridiag(A):
    n = A.shape[0]
    a = np.asfortranarray(A, dtype=np.float64)
    d, s = np.zeros(n), np.zeros(n - 1)
    _ck(lib().orc_tridiag(n, a.ctypes.data_as(C.POINTER(C.c_double)), _p(d, C.c_double), _p(s, C.c_double)))
    return d, s


def tridiag_eig(d, s):
    d = np.ascontiguousarray(d, dtype=np.float64)
    s = np.ascontiguousarray(s, dtype=np.float64)
    ev = np.zeros(len(d))
    _ck(lib().orc_tridiag_eig(len(d), _p(d, C.c_double), _p(s, C.c_double), _p(ev, C.c_double)))
    return ev


def calc_ed(kind, L, f, U, mu_c, beta, t=1.0, tp=1.0, vectors=False):
    n = lattice_size(kind, L)
    f = np.ascontiguousarray(f, dtype=np.int32)
    sp, ex, fe = np.zeros(n), np.zeros(n), np.zeros(n)
    vec = np.zeros((n, n), order="F") if vectors else None
    logz = C.c_double(0)
    _ck(lib().orc_calc_ed(kind, L, C.c_double(t), C.c_double(tp), _p(f, C.c_int), C.c_double(U), C.c_double(mu_c),
                          C.c_double(beta), int(vectors), _p(sp, C.c_double), _p(ex, C.c_double), _p(fe, C.c_double),
                          None if vec is None else vec.ctypes.data_as(C.POINTER(C.c_double)), C.byref(logz)))
    return dict(spectrum=sp, cached_exp=ex, cached_fermi=fe, evecs=vec, logZ=logz.value)


def logz_from_spectrum(spectrum, beta):
    s = np.ascontiguousarray(spectrum, dtype=np.float64)
    return lib().orc_logz_from_spectrum(len(s), _p(s, C.c_double), C.c_double(beta))


def cheb_sizes(msize, prefactor):
    M, G = C.c_int(0), C.c_int(0)
    lib().orc_cheb_sizes(msize, C.c_double(prefactor), C.byref(M), C.byref(G))
    return M.value, G.value


def cheb_table(M, G):
    Me = M + M % 2
    T, x, th = np.zeros((Me, G)), np.zeros(G), np.zeros(G)
    _ck(lib().orc_cheb_table(M, G, _p(T, C.c_double), _p(x, C.c_double), _p(th, C.c_double)))
    return T, x, th


def cheb_moment(M, G, vals, order):
    v = np.ascontiguousarray(vals, dtype=np.float64)
    return lib().orc_cheb_moment(M, G, _p(v, C.c_double), order)


def calc_chebyshev(kind, L, f, U, mu_c, beta, M, G, t=1.0, tp=1.0, emode=0, prune=False):
    f = np.ascontiguousarray(f, dtype=np.int32)
    mom, out5 = np.zeros(M + M % 2), np.zeros(5)
    steps = C.c_int(0)
    _ck(lib().orc_calc_chebyshev(kind, L, C.c_double(t), C.c_double(tp), _p(f, C.c_int), C.c_double(U), C.c_double(mu_c),
                                 C.c_double(beta), M, G, emode, int(prune), _p(mom, C.c_double), _p(out5, C.c_double),
                                 C.byref(steps)))
    return dict(moments=mom, e_min=out5[0], e_max=out5[1], a=out5[2], b=out5[3], logZ=out5[4], lanczos_steps=steps.value)


def measure_ipr(evecs):
    n = evecs.shape[0]
    v = np.asfortranarray(evecs, dtype=np.float64)
    out = np.zeros(n)
    _ck(lib().orc_measure_ipr(n, v.ctypes.data_as(C.POINTER(C.c_double)), _p(out, C.c_double)))
    return out


def make_params(kind=CUBIC2D, L=8, t=1.0, tp=1.0, beta=1.0, U=1.0, mu_c=None, mu_f=None, mc_flip=0.0, mc_add_remove=1.0,
                mc_reshuffle=0.0, cheb_moves=False, cheb_prefactor=2.2, emode=0, seed=32167, nf_start=None, nsweeps=8,
                sweep_len=16, ntherm_sweeps=1, measure_energy=True, measure_ipr=False, W=()):
    n = lattice_size(kind, L)
    W = [float(w) for w in W]
    return McParams(kind, L, t, tp, beta, U, U / 2 if mu_c is None else mu_c, U / 2 if mu_f is None else mu_f, mc_flip,
                    mc_add_remove, mc_reshuffle, int(cheb_moves), cheb_prefactor, emode, seed,
                    n // 2 if nf_start is None else nf_start, nsweeps, sweep_len, ntherm_sweeps, int(measure_energy),
                    int(measure_ipr), len(W), (C.c_double * 8)(*(W + [0.0] * (8 - len(W)))))


def mc_run(p, rank=0, trace=True):
    n = lattice_size(p.kind, p.L)
    steps = (p.nsweeps + p.ntherm_sweeps) * p.sweep_len
    tr = dict(move=np.zeros(steps, np.int32), site_a=np.zeros(steps, np.int32), site_b=np.zeros(steps, np.int32),
              accepted=np.zeros(steps, np.int32), weight=np.zeros(steps), u=np.zeros(steps), logz_new=np.zeros(steps))
    en, d2, ce, sp = np.zeros(p.nsweeps), np.zeros(p.nsweeps), np.zeros(p.nsweeps), np.zeros(n)
    f_final = np.zeros(n, np.int32)
    nacc, lzf = C.c_long(0), C.c_double(0)
    ipr = np.zeros((p.nsweeps, n)) if p.measure_ipr else None
    ti = (lambda k: _p(tr[k], C.c_int)) if trace else (lambda k: None)
    td = (lambda k: _p(tr[k], C.c_double)) if trace else (lambda k: None)
    _ck(lib().orc_mc_run(C.byref(p), rank, ti("move"), ti("site_a"), ti("site_b"), ti("accepted"), td("weight"), td("u"),
                         td("logz_new"), _p(en, C.c_double), _p(d2, C.c_double), _p(ce, C.c_double), _p(sp, C.c_double),
                         _p(f_final, C.c_int), C.byref(nacc), C.byref(lzf), _p(ipr, C.c_double)))
    return dict(trace=tr if trace else None, energies=en, d2energies=d2, c_energies=ce, spectrum_avg=sp, f_final=f_final,
                naccept=nacc.value, logz_final=lzf.value, ipr_history=ipr)


def mc_histories(p, rank=0):
    """spectrum_history [nsweeps, N] and focc_history [nsweeps, V] of one oracle chain."""
    n = lattice_size(p.kind, p.L)
    sh, fo = np.zeros((p.nsweeps, n)), np.zeros((p.nsweeps, n), np.int32)
    _ck(lib().orc_mc_histories(C.byref(p), rank, _p(sh, C.c_double), _p(fo, C.c_int)))
    return sh, fo


def bench_chains(p, nthreads, rank0=0):
    sec, nacc = C.c_double(0), C.c_long(0)
    _ck(lib().orc_bench_chains(C.byref(p), nthreads, rank0, C.byref(sec), C.byref(nacc)))
    return sec.value, nacc.value


def binning(x, max_depth):
    x = np.ascontiguousarray(x, dtype=np.float64)
    rows = np.zeros((max_depth + 1, 5))
    if lib().orc_binning(len(x), _p(x, C.c_double), max_depth, _p(rows, C.c_double)) != 0:
        raise RuntimeError("oracle binning failed")
    return rows


def jackknife(series, depth):
    d = np.ascontiguousarray(series, dtype=np.float64)
    if d.ndim == 1:
        d = d[None]
    out = np.zeros(4)
    if lib().orc_jackknife(d.shape[1], d.shape[0], _p(d, C.c_double), depth, _p(out, C.c_double)) != 0:
        raise RuntimeError("oracle jackknife failed")
    return out


def stiffness(kind, L, f, U, mu_c, beta, offset=0.05, wgrid=(0.0,), t=1.0):
    f = np.ascontiguousarray(f, dtype=np.int32)
    wg = np.ascontiguousarray(wgrid, dtype=np.float64)
    cond = np.zeros(len(wg))
    st = C.c_double(0)
    if lib().orc_stiffness(kind, L, C.c_double(t), _p(f, C.c_int), C.c_double(U), C.c_double(mu_c), C.c_double(beta), C.c_double(offset),
                           len(wg), _p(wg, C.c_double), _p(cond, C.c_double), C.byref(st)) != 0:
        raise RuntimeError("oracle stiffness failed")
    return st.value, cond
