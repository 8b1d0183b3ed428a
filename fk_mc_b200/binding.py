"""ctypes binding of libfkmc_b200.so (include/fkmc.h).  No torch types cross this boundary."""
import ctypes as C
import math
import os
import re
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.environ.get("FKMC_LIB", os.path.join(PKG_DIR, "lib", "libfkmc_b200.so"))  # FKMC_LIB: developer builds (instrumented kernels)
HEADER = os.path.join(ROOT, "include", "fkmc.h")

KINDS = {"cubic1d": 1, "cubic2d": 2, "cubic3d": 3, "triangular": 4, "honeycomb": 5, "honeycomb_ref_lower": 7}
MOVE_FLIP, MOVE_ADDREMOVE, MOVE_RESHUFFLE = 0, 1, 2

STATUS = {0: "OK", 1: "INVALID", 2: "CUDA", 3: "NO_DEVICE", 4: "NOCONV", 5: "STATE"}


class FkmcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("fkmc error %s: %s" % (STATUS.get(code, code), msg))
        self.code = code


class ChainParams(C.Structure):
    """fkmc_chain_params (include/fkmc.h); mirrors fk_mc<L>::define_parameters (fk_mc.hxx:177-207)."""
    _fields_ = [("beta", C.c_double), ("U", C.c_double), ("mu_c", C.c_double), ("mu_f", C.c_double),
                ("mc_flip", C.c_double), ("mc_add_remove", C.c_double), ("mc_reshuffle", C.c_double),
                ("cheb_moves", C.c_int32), ("cheb_prefactor", C.c_double), ("seed", C.c_int64), ("chain0", C.c_int32),
                ("nf_start", C.c_int32), ("sweep_len", C.c_int32), ("ntherm_sweeps", C.c_int32),
                ("measure_energy", C.c_int32), ("record_trace", C.c_int32), ("max_sweeps", C.c_int32),
                ("measure_history", C.c_int32), ("measure_ipr", C.c_int32), ("n_W", C.c_int32), ("W", C.c_double * 8),
                ("measure_eigenfunctions", C.c_int32), ("measure_stiffness", C.c_int32), ("n_cond_w", C.c_int32), ("cond_offset", C.c_double),
                ("cond_wgrid", C.c_double * 32), ("fast_update", C.c_int32), ("fu_refresh_sweeps", C.c_int32)]


def build_library(force=False):
    """Compile every CUDA source for sm_100a into fk_mc_b200/lib/libfkmc_b200.so (nvcc cross-compiles without a GPU)."""
    src = os.path.join(PKG_DIR, "csrc")
    if force:
        subprocess.check_call(["make", "-C", src, "clean"])
    subprocess.check_call(["make", "-C", src, "-j8", "-s"])
    return LIB_PATH


def exported_symbols():
    """Function names declared in include/fkmc.h."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fkmc_[a-z0-9_]+)\s*\(", text)))


_lib = None


def load_library():
    """Load the C-ABI library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FkmcError(3, "libfkmc_b200.so is missing (run __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    lib.fkmc_last_error.restype = C.c_char_p
    lib.fkmc_last_error.argtypes = [C.c_void_p]
    lib.fkmc_launch_count.restype = C.c_int64
    lib.fkmc_launch_count.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def nccl_unique_id():
    """128-byte ncclUniqueId (rank 0 calls this and hands the bytes to the other ranks over any host channel)."""
    buf = (C.c_char * 128)()
    rc = load_library().fkmc_nccl_unique_id(buf)
    if rc != 0:
        raise FkmcError(rc, load_library().fkmc_last_error(None).decode())
    return bytes(buf.raw)


def cheb_sizes(msize, prefactor=2.2):
    """fk_mc.hxx:60-63: M = even(int(ln N * prefactor)), G = max(2M, 10)."""
    m = int(math.log(float(msize)) * prefactor)
    m += m % 2
    return m, max(2 * m, 10)


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class Context:
    """One GPU's fkmc_ctx.  Lattice == hypercubic_lattice<D>(L) + fill_* of the reference."""

    def __init__(self, kind, L, t=1.0, tp=1.0, max_batch=1, device=0):
        self.lib = load_library()
        self.kind = KINDS[kind] if isinstance(kind, str) else int(kind)
        self.h = C.c_void_p()
        rc = self.lib.fkmc_create(C.byref(self.h), int(device), self.kind, int(L), C.c_double(t), C.c_double(tp),
                                  int(max_batch))
        if rc != 0:
            raise FkmcError(rc, self.lib.fkmc_last_error(None).decode())
        self.L = L
        self.N = self.lib.fkmc_volume(self.h)
        self.max_batch = max_batch
        self.chain_params = None
        self.n_chains = 0

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.fkmc_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise FkmcError(rc, self.lib.fkmc_last_error(self.h).decode())

    # ---- plumbing ----
    def set_stream(self, cuda_stream):
        self._ck(self.lib.fkmc_set_stream(self.h, C.c_void_p(cuda_stream)))

    def sync(self):
        self._ck(self.lib.fkmc_sync(self.h))

    def launch_count(self):
        return int(self.lib.fkmc_launch_count(self.h))

    def hopping_dense(self):
        H = np.zeros((self.N, self.N))
        self._ck(self.lib.fkmc_hopping_dense(self.h, _ptr(H, C.c_double)))
        return H

    def timer_begin(self):
        self._ck(self.lib.fkmc_timer_begin(self.h))

    def timer_end(self):
        ms = C.c_float(0)
        self._ck(self.lib.fkmc_timer_end(self.h, C.byref(ms)))
        return ms.value

    def profile_enable(self, on=True):
        self._ck(self.lib.fkmc_profile_enable(self.h, int(on)))

    def profile_reset(self):
        self._ck(self.lib.fkmc_profile_reset(self.h))

    def profile_get(self, family):
        ms, n = C.c_double(0), C.c_int64(0)
        self._ck(self.lib.fkmc_profile_get(self.h, family.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # ---- weight evaluators ----
    def _f(self, f):
        f = np.ascontiguousarray(f, dtype=np.int32)
        if f.ndim == 1:
            f = f[None, :]
        if f.shape[1] != self.N:
            raise FkmcError(1, "f must have shape [B, V]")
        return f

    @staticmethod
    def _out(out, key, shape):
        """Result buffer: the caller's (e.g. page-locked, reused from call to call: the device -> host copies then run at full PCIe rate
        and nothing is allocated per call) or a fresh array."""
        if out is None or key not in out:
            return np.zeros(shape)
        a = out[key]
        if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.shape == tuple(shape)):
            raise FkmcError(1, "out[%r] must be a C-contiguous float64 array of shape %r" % (key, tuple(shape)))
        return a

    def logz_ed(self, f, U, mu_c, beta, want_caches=False, out=None):
        """configuration_t::calc_ed(false): returns dict(spectrum [B,N], logZ [B], cached_exp, cached_fermi).
        out: optional dict of preallocated result arrays (spectrum, logZ)."""
        f = self._f(f)
        B = f.shape[0]
        ev, lz = self._out(out, "spectrum", (B, self.N)), self._out(out, "logZ", (B,))
        ex = np.zeros((B, self.N)) if want_caches else None
        fe = np.zeros((B, self.N)) if want_caches else None
        self._ck(self.lib.fkmc_logz_ed_batched(self.h, _ptr(f, C.c_int32), B, C.c_double(U), C.c_double(mu_c), C.c_double(beta),
                                               _ptr(ev, C.c_double), _ptr(lz, C.c_double), _ptr(ex, C.c_double),
                                               _ptr(fe, C.c_double)))
        return dict(spectrum=ev, logZ=lz, cached_exp=ex, cached_fermi=fe)

    def eigh(self, f, U, mu_c, beta):
        """configuration_t::calc_ed(true): dict(spectrum [B,N], evecs [B,N,N] with evecs[b][:, k] <-> spectrum[b][k], logZ [B])."""
        f = self._f(f)
        B = f.shape[0]
        ev, lz = np.zeros((B, self.N)), np.zeros(B)
        vec = np.zeros((B, self.N, self.N))  # column-major per matrix == [k][i]
        self._ck(self.lib.fkmc_eigh_batched(self.h, _ptr(f, C.c_int32), B, C.c_double(U), C.c_double(mu_c), C.c_double(beta),
                                            _ptr(ev, C.c_double), _ptr(vec, C.c_double), _ptr(lz, C.c_double)))
        return dict(spectrum=ev, evecs=np.transpose(vec, (0, 2, 1)), logZ=lz)

    def ipr(self, f, U, mu_c, beta):
        """measure_ipr::accumulate: dict(spectrum [B,N], ipr [B,N])."""
        f = self._f(f)
        B = f.shape[0]
        ev, ip = np.zeros((B, self.N)), np.zeros((B, self.N))
        self._ck(self.lib.fkmc_ipr_batched(self.h, _ptr(f, C.c_int32), B, C.c_double(U), C.c_double(mu_c), C.c_double(beta),
                                           _ptr(ev, C.c_double), _ptr(ip, C.c_double)))
        return dict(spectrum=ev, ipr=ip)

    def stiffness(self, f, U, mu_c, beta, offset=0.05, wgrid=(0.0,)):
        """measure_stiffness::accumulate on the GPU: returns (stiffness [B], conductivity [B, n_w])."""
        f = self._f(f)
        B = f.shape[0]
        wg = np.ascontiguousarray(wgrid, dtype=np.float64)
        st, cd = np.zeros(B), np.zeros((B, max(len(wg), 1)))
        self._ck(self.lib.fkmc_stiffness_batched(self.h, _ptr(f, C.c_int32), B, C.c_double(U), C.c_double(mu_c), C.c_double(beta),
                                                 C.c_double(offset), len(wg), _ptr(wg, C.c_double), _ptr(st, C.c_double), _ptr(cd, C.c_double)))
        return st, cd[:, :len(wg)]

    def chain_ipr(self):
        ev, ip = np.zeros((self.n_chains, self.N)), np.zeros((self.n_chains, self.N))
        self._ck(self.lib.fkmc_chain_ipr(self.h, _ptr(ev, C.c_double), _ptr(ip, C.c_double)))
        return dict(spectrum=ev, ipr=ip)

    def logz_kpm(self, f, U, mu_c, beta, M, G):
        """configuration_t::calc_chebyshev: returns dict(moments [B,M], e_min, e_max, a, b, logZ [B])."""
        f = self._f(f)
        B = f.shape[0]
        mom, ab, lz = np.zeros((B, M)), np.zeros((B, 4)), np.zeros(B)
        self._ck(self.lib.fkmc_logz_kpm_batched(self.h, _ptr(f, C.c_int32), B, C.c_double(U), C.c_double(mu_c), C.c_double(beta),
                                                int(M), int(G), _ptr(mom, C.c_double), _ptr(ab, C.c_double),
                                                _ptr(lz, C.c_double)))
        return dict(moments=mom, e_min=ab[:, 0], e_max=ab[:, 1], a=ab[:, 2], b=ab[:, 3], logZ=lz)

    def logz_kpm_local(self, f, U, mu_c, beta, M, G, f_ref=None, state_ref=None, out=None):
        """calc_chebyshev for configurations that differ from f_ref in one or two sites (fkmc_logz_kpm_batched_local); returns the dict of
        logz_kpm plus state [B, 64], the record to pass as state_ref when f becomes the reference.
        out: optional dict of preallocated result arrays (moments, ab [B, 4], logZ, state)."""
        f = self._f(f)
        B = f.shape[0]
        mom, ab, lz, st = self._out(out, "moments", (B, M)), self._out(out, "ab", (B, 4)), self._out(out, "logZ", (B,)), self._out(out, "state", (B, 64))
        fr = sr = None
        if f_ref is not None:
            fr = self._f(f_ref)
            sr = np.ascontiguousarray(state_ref, dtype=np.float64).reshape(B, 64)
        self._ck(self.lib.fkmc_logz_kpm_batched_local(self.h, _ptr(f, C.c_int32), _ptr(fr, C.c_int32) if fr is not None else None,
                                                      _ptr(sr, C.c_double) if sr is not None else None, B, C.c_double(U), C.c_double(mu_c),
                                                      C.c_double(beta), int(M), int(G), _ptr(mom, C.c_double), _ptr(ab, C.c_double),
                                                      _ptr(lz, C.c_double), _ptr(st, C.c_double)))
        return dict(moments=mom, e_min=ab[:, 0], e_max=ab[:, 1], a=ab[:, 2], b=ab[:, 3], logZ=lz, state=st)

    def energy_from_spectrum(self, evals, beta):
        ev = np.ascontiguousarray(evals, dtype=np.float64)
        if ev.ndim == 1:
            ev = ev[None, :]
        out = np.zeros((ev.shape[0], 3))
        self._ck(self.lib.fkmc_energy_from_spectrum(self.h, _ptr(ev, C.c_double), ev.shape[0], C.c_double(beta),
                                                    _ptr(out, C.c_double)))
        return out

    # ---- stage-level ----
    def sytrd(self, A):
        """A: [B, N, N] symmetric matrices (the lower triangle is read).  Returns d [B,N], e [B,N-1]."""
        A = np.asarray(A, dtype=np.float64)
        if A.ndim == 2:
            A = A[None]
        B, n, _ = A.shape
        Af = np.ascontiguousarray(np.transpose(A, (0, 2, 1)))  # column-major per matrix
        d, e = np.zeros((B, n)), np.zeros((B, n - 1))
        self._ck(self.lib.fkmc_sytrd_batched(self.h, _ptr(Af, C.c_double), n, B, _ptr(d, C.c_double), _ptr(e, C.c_double)))
        return d, e

    def sy2sb(self, A):
        """Stage 1 of the two-stage reduction: [B, N, N] symmetric (lower read) -> band storage [B, 9, N]."""
        A = np.asarray(A, dtype=np.float64)
        if A.ndim == 2:
            A = A[None]
        B, n, _ = A.shape
        Af = np.ascontiguousarray(np.transpose(A, (0, 2, 1)))
        AB = np.zeros((B, 9, n))
        self._ck(self.lib.fkmc_sy2sb_batched(self.h, _ptr(Af, C.c_double), n, B, _ptr(AB, C.c_double)))
        return AB

    def sb2st(self, AB):
        """Stage 2: band storage [B, 9, N] -> tridiagonal d [B, N], e [B, N-1]."""
        AB = np.ascontiguousarray(AB, dtype=np.float64)
        if AB.ndim == 2:
            AB = AB[None]
        B, _, n = AB.shape
        d, e = np.zeros((B, n)), np.zeros((B, n - 1))
        self._ck(self.lib.fkmc_sb2st_batched(self.h, _ptr(AB, C.c_double), n, B, _ptr(d, C.c_double), _ptr(e, C.c_double)))
        return d, e

    def set_option(self, name, value):
        self._ck(self.lib.fkmc_set_option(self.h, name.encode(), int(value)))

    def tridiag_eigvals(self, d, e):
        d = np.ascontiguousarray(d, dtype=np.float64)
        e = np.ascontiguousarray(e, dtype=np.float64)
        if d.ndim == 1:
            d, e = d[None], e[None]
        B, n = d.shape
        ev = np.zeros((B, n))
        self._ck(self.lib.fkmc_tridiag_eigvals_batched(self.h, _ptr(d, C.c_double), _ptr(e, C.c_double), n, B,
                                                       _ptr(ev, C.c_double)))
        return ev

    def secular_update(self, lam, z, rho):
        """Eigenvalues of diag(lam) + rho z z^T (batched rank-one secular solver): lam, z [B, N], rho [B] -> [B, N]."""
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        z = np.ascontiguousarray(z, dtype=np.float64)
        if lam.ndim == 1:
            lam, z = lam[None], z[None]
        B, n = lam.shape
        rho = np.ascontiguousarray(np.broadcast_to(np.asarray(rho, dtype=np.float64), (B,)))
        out = np.zeros((B, n))
        self._ck(self.lib.fkmc_secular_update_batched(self.h, _ptr(lam, C.c_double), _ptr(z, C.c_double), _ptr(rho, C.c_double), n, B,
                                                      _ptr(out, C.c_double)))
        return out

    def rng_stream(self, seed, mode, V, count):
        out = np.zeros(count)
        self._ck(self.lib.fkmc_rng_stream(self.h, C.c_int64(seed), mode, V, count, _ptr(out, C.c_double)))
        return out

    # ---- chains ----
    def chain_init(self, n_chains, beta, U, mu_c=None, mu_f=None, mc_flip=0.0, mc_add_remove=1.0, mc_reshuffle=0.0,
                   cheb_moves=False, cheb_prefactor=2.2, seed=32167, chain0=0, nf_start=None, sweep_len=16, ntherm_sweeps=1,
                   measure_energy=True, record_trace=False, max_sweeps=64, measure_history=False, measure_ipr=False, W=(),
                   fast_update=False, fu_refresh_sweeps=0, measure_eigenfunctions=False, measure_stiffness=False, cond_wgrid=(), cond_offset=0.05):
        cw = [float(w) for w in cond_wgrid]
        if len(cw) > 32:
            raise FkmcError(1, "at most 32 conductivity frequencies")
        W = [float(w) for w in W]
        if len(W) > 8:
            raise FkmcError(1, "at most 8 f-f interaction terms")
        p = ChainParams(beta, U, U / 2 if mu_c is None else mu_c, U / 2 if mu_f is None else mu_f, mc_flip, mc_add_remove,
                        mc_reshuffle, int(cheb_moves), cheb_prefactor, seed, chain0,
                        self.N // 2 if nf_start is None else nf_start, sweep_len, ntherm_sweeps, int(measure_energy),
                        int(record_trace), max_sweeps, int(measure_history), int(measure_ipr), len(W),
                        (C.c_double * 8)(*(W + [0.0] * (8 - len(W)))), int(measure_eigenfunctions), int(measure_stiffness), len(cw),
                        float(cond_offset), (C.c_double * 32)(*(cw + [0.0] * (32 - len(cw)))), int(fast_update), int(fu_refresh_sweeps))
        self._ck(self.lib.fkmc_chain_init(self.h, int(n_chains), C.byref(p)))
        self.chain_params = p
        self.n_chains = n_chains
        return p

    def chain_run_sweeps(self, n_sweeps):
        self._ck(self.lib.fkmc_chain_run_sweeps(self.h, int(n_sweeps)))

    def chain_get_series(self):
        p, Cn = self.chain_params, self.n_chains
        e = np.zeros((p.max_sweeps, Cn))
        d2, ec = np.zeros_like(e), np.zeros_like(e)
        nf = np.zeros((p.max_sweeps, Cn), dtype=np.int32)
        n = C.c_int(0)
        self._ck(self.lib.fkmc_chain_get_series(self.h, C.byref(n), _ptr(e, C.c_double), _ptr(d2, C.c_double),
                                                _ptr(ec, C.c_double), _ptr(nf, C.c_int32)))
        m = n.value
        return dict(n_measured=m, energies=e[:m], d2energies=d2[:m], c_energies=ec[:m], nf=nf[:m])

    def chain_get_history(self):
        """Per-sweep histories: dict(n_measured, spectrum_mean [C,N], spectrum_history / focc_history / ipr_history [n,C,N]);
        entries whose measure is off are None."""
        p, Cn, N = self.chain_params, self.n_chains, self.N
        exact = (not p.cheb_moves) or p.measure_energy or p.measure_ipr or p.measure_eigenfunctions
        sm = np.zeros((Cn, N)) if exact else None
        sh = np.zeros((p.max_sweeps, Cn, N)) if exact and p.measure_history else None
        fo = np.zeros((p.max_sweeps, Cn, N), dtype=np.int32) if p.measure_history else None
        ip = np.zeros((p.max_sweeps, Cn, N)) if p.measure_ipr else None
        n = C.c_int(0)
        self._ck(self.lib.fkmc_chain_get_history(self.h, C.byref(n), _ptr(sm, C.c_double), _ptr(sh, C.c_double), _ptr(fo, C.c_int32),
                                                 _ptr(ip, C.c_double)))
        m = n.value
        cut = lambda a: None if a is None else a[:m]  # noqa: E731
        return dict(n_measured=m, spectrum_mean=sm, spectrum_history=cut(sh), focc_history=cut(fo), ipr_history=cut(ip))

    def chain_get_eigenfunctions(self):
        """eigenfunctions_history: [n_measured, n_chains, N, N] with [..., i, k] = component i of eigenvector k."""
        p, Cn, N = self.chain_params, self.n_chains, self.N
        ev = np.zeros((p.max_sweeps, Cn, N, N))
        n = C.c_int(0)
        self._ck(self.lib.fkmc_chain_get_eigenfunctions(self.h, C.byref(n), _ptr(ev, C.c_double)))
        return np.transpose(ev[:n.value], (0, 1, 3, 2))   # stored eigenvector-major (column-major matrix)

    def chain_get_stiffness(self):
        """dict(n_measured, stiffness [n_measured, n_chains], cond [n_measured, n_chains, n_cond_w])."""
        p, Cn = self.chain_params, self.n_chains
        nw = p.n_cond_w
        st = np.zeros((p.max_sweeps, Cn))
        cd = np.zeros((p.max_sweeps, Cn, max(nw, 1)))
        n = C.c_int(0)
        self._ck(self.lib.fkmc_chain_get_stiffness(self.h, C.byref(n), _ptr(st, C.c_double), _ptr(cd, C.c_double) if nw else None))
        m = n.value
        return dict(n_measured=m, stiffness=st[:m], cond=cd.reshape(-1)[: m * Cn * nw].reshape(m, Cn, nw))

    def chain_get_fsector(self):
        """dict(n_measured, nf0, nfpi [n_measured, n_chains]): n_f(q=0) and |n_f(q=pi)| per measured sweep (fsusc0pi.hpp:36-46)."""
        p, Cn = self.chain_params, self.n_chains
        n0 = np.zeros((p.max_sweeps, Cn), dtype=np.int32)
        npi = np.zeros((p.max_sweeps, Cn), dtype=np.int32)
        n = C.c_int(0)
        has0 = bool(p.measure_energy) or not p.cheb_moves
        self._ck(self.lib.fkmc_chain_get_fsector(self.h, C.byref(n), _ptr(n0, C.c_int32) if has0 else None, _ptr(npi, C.c_int32)))
        m = n.value
        return dict(n_measured=m, nf0=n0[:m] if has0 else None, nfpi=npi[:m])

    def chain_get_state(self, spectrum=False):
        Cn = self.n_chains
        f = np.zeros((Cn, self.N), dtype=np.int32)
        lz = np.zeros(Cn)
        na = np.zeros(Cn, dtype=np.int64)
        sp = np.zeros((Cn, self.N)) if spectrum else None
        self._ck(self.lib.fkmc_chain_get_state(self.h, _ptr(f, C.c_int32), _ptr(lz, C.c_double), _ptr(na, C.c_int64),
                                               _ptr(sp, C.c_double)))
        return dict(f=f, logZ=lz, naccept=na, spectrum=sp)

    def chain_get_trace(self):
        p, Cn = self.chain_params, self.n_chains
        steps = p.max_sweeps * p.sweep_len
        ti = lambda: np.zeros((steps, Cn), dtype=np.int32)  # noqa: E731
        td = lambda: np.zeros((steps, Cn))  # noqa: E731
        move, a, b, acc, w, u, lz = ti(), ti(), ti(), ti(), td(), td(), td()
        n = C.c_int(0)
        self._ck(self.lib.fkmc_chain_get_trace(self.h, C.byref(n), _ptr(move, C.c_int32), _ptr(a, C.c_int32), _ptr(b, C.c_int32),
                                               _ptr(acc, C.c_int32), _ptr(w, C.c_double), _ptr(u, C.c_double),
                                               _ptr(lz, C.c_double)))
        k = n.value
        return dict(n_steps=k, move=move[:k], site_a=a[:k], site_b=b[:k], accepted=acc[:k], weight=w[:k], u=u[:k],
                    logz_new=lz[:k])

    # ---- end-of-run collective (NCCL resolved inside the library; no torch involved) ----
    def comm_init(self, unique_id, nranks, rank):
        """Join the NCCL communicator described by the 128-byte id rank 0 obtained from nccl_unique_id() (collective)."""
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        self._ck(self.lib.fkmc_comm_init(self.h, buf, int(nranks), int(rank)))
        self.nranks = nranks

    def gather_series(self):
        """All-gather of the energy series over every rank's chains: dict(n_measured, energies / d2energies / c_energies
        [n_measured, nranks * n_chains] in global chain order).  Local series when no communicator was initialised."""
        nr = getattr(self, "nranks", 1)
        p, Cn = self.chain_params, self.n_chains
        e = np.zeros((p.max_sweeps, nr * Cn))
        d2, ec = np.zeros_like(e), np.zeros_like(e)
        n, tot = C.c_int(0), C.c_int(0)
        self._ck(self.lib.fkmc_gather_series(self.h, C.byref(n), C.byref(tot), _ptr(e, C.c_double), _ptr(d2, C.c_double),
                                             _ptr(ec, C.c_double), None))
        m = n.value
        flat = lambda a: a.reshape(-1)[: m * tot.value].reshape(m, tot.value)  # noqa: E731
        return dict(n_measured=m, energies=flat(e), d2energies=flat(d2), c_energies=flat(ec))

    def kpm_last_steps(self, B):
        """Lanczos steps each of the last B KPM evaluations needed (the ARPACK iteration-count analogue)."""
        st = np.zeros(B, dtype=np.int32)
        self._ck(self.lib.fkmc_kpm_last_steps(self.h, int(B), _ptr(st, C.c_int32)))
        return st

    def chain_series_dev(self):
        e, d2, ec, ld = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int(0)
        self._ck(self.lib.fkmc_chain_series_dev(self.h, C.byref(e), C.byref(d2), C.byref(ec), C.byref(ld)))
        return e.value, d2.value, ec.value, ld.value
