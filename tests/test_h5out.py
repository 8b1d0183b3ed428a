"""HDF5 output writer (fk_mc_b200/h5out.py): the reader is pinned on a file written by the real HDF5 library, the writer
is checked through that reader, and save_all_data against the reference layout (prog/data_save.hxx:33-151)."""
import glob
import os
import struct

import numpy as np
import pytest

from fk_mc_b200 import h5out, stats


def _real_hdf5_fixture():
    import scipy.io
    hits = glob.glob(os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat"))
    return hits[0] if hits else None


def test_reader_parses_a_file_written_by_libhdf5():
    fn = _real_hdf5_fixture()
    if fn is None:
        pytest.skip("scipy's HDF5 fixture is not installed")
    r = h5out.H5Reader(fn)
    assert r.base == 512 and (r.leaf_k, r.internal_k) == (4, 16)          # MATLAB user block, library default K values
    t = r.tree()
    assert list(t) == ["/testdouble"]
    assert np.allclose(t["/testdouble"].ravel(), np.arange(9) * np.pi / 4)  # the fixture's content (scipy test_mio.py)


def test_writer_round_trip_and_structure(tmp_path):
    w = h5out.H5Writer()
    rng = np.random.default_rng(0)
    big = rng.standard_normal((7, 33))
    w["/mc_data/energies"] = np.arange(10.0)
    w["/mc_data/ipr_history"] = big
    w["/mc_data/empty"] = np.zeros(0)
    w["/stats/energy"] = np.array([50, -0.2474323, 1.207943e-28, 1.554312e-15])
    w["/parameters/beta"] = 10.0
    w["/parameters/L"] = 32
    w["/parameters/seed"] = np.int64(2 ** 40 + 3)
    w["/parameters/cheb_moves"] = True
    w["/parameters/output"] = "output.h5"
    for i in range(40):                       # more members than one symbol-table node holds (2 * LEAF_K = 8)
        w["/many/p%02d" % i] = float(i)
    fn = str(tmp_path / "t.h5")
    size = w.save(fn)
    raw = open(fn, "rb").read()
    assert len(raw) == size and raw[:8] == b"\x89HDF\r\n\x1a\n"
    assert struct.unpack("<Q", raw[40:48])[0] == size                      # end-of-file address in the superblock
    r = h5out.H5Reader(fn)
    t = r.tree()
    assert np.array_equal(t["/mc_data/energies"], np.arange(10.0)) and np.array_equal(t["/mc_data/ipr_history"], big)
    assert t["/mc_data/empty"].shape == (0,)
    assert t["/parameters/beta"] == 10.0 and t["/parameters/L"] == 32 and t["/parameters/seed"] == 2 ** 40 + 3
    assert t["/parameters/cheb_moves"] == 1 and t["/parameters/output"] == "output.h5"
    assert [t["/many/p%02d" % i] for i in range(40)] == [float(i) for i in range(40)]
    assert sorted(r["/"]) == ["many", "mc_data", "parameters", "stats"]
    # the float64 datatype message is byte-identical to the one libhdf5 writes (taken from the fixture above)
    fx = _real_hdf5_fixture()
    if fx:
        rr = h5out.H5Reader(fx)
        theirs = [d for t_, d in rr._messages(rr.members()["testdouble"]) if t_ == 3][0]
        ours = [d for t_, d in r._messages(r["/mc_data"]["energies"]) if t_ == 3][0]
        assert ours == theirs
    with pytest.raises(KeyError):
        r["/nope"]


def test_save_all_data_layout(tmp_path):
    rng = np.random.default_rng(1)
    n, beta, vol = 256, 10.0, 64
    e = -0.25 + 0.01 * rng.standard_normal(n)
    d2 = 0.02 + 0.001 * rng.standard_normal(n)
    ce = e + 0.5
    fn = str(tmp_path / "output.h5")
    params = dict(beta=beta, U=1.0, L=8, mu_c=0.5, mu_f=0.5, nsweeps=n, sweep_len=16, cheb_moves=False, output="output.h5")
    out = h5out.save_all_data(fn, params, e, d2, ce, beta, vol, histories={"ipr_history": rng.random((64, 4))})
    t = h5out.H5Reader(fn).tree()
    assert {"/mc_data/energies", "/mc_data/d2energies", "/mc_data/c_energies", "/mc_data/ipr_history", "/parameters/beta",
            "/stats/energy", "/stats/d2energy", "/stats/c_energy", "/stats/cv", "/binning/energy", "/binning/cv"} <= set(t)
    assert np.array_equal(t["/mc_data/energies"], e)
    rep = stats.energy_report(e, d2, beta, vol)
    assert t["/stats/energy"].shape == (4,) and np.allclose(t["/stats/energy"], rep["energy"]["stats"])
    assert np.allclose(t["/stats/cv"], rep["cv"]["stats"])
    assert t["/binning/energy"].shape == (len(rep["energy"]["binning"]), 5)          # [n, mean, variance, stderr, tau_int]
    assert np.allclose(t["/binning/energy"][:, 4], rep["energy"]["cor_length"])
    assert np.allclose(out["cv"][1], t["/stats/cv"])
    # consumer convention (scripts/parse/parse_thermod.py:48-49): (nbins, value, disp, error) = h5["stats"][obs]
    nb, val, disp, err = t["/stats/energy"]
    assert nb >= 4 and abs(val - e.mean()) < 1e-12 and err > 0


def test_save_all_data_multi_chain_and_fsector(tmp_path):
    """[measurement][chain] series are written chain after chain (the reference's gathered ranks) and the f-sector statistics of
    save_fstats (prog/data_save.hxx:200-236) land under /stats and /binning."""
    rng = np.random.default_rng(2)
    n_meas, n_chains = 64, 8
    e = rng.standard_normal((n_meas, n_chains))
    nf0 = rng.integers(20, 44, size=(n_meas, n_chains))
    nfpi = np.abs(rng.integers(-9, 9, size=(n_meas, n_chains)))
    fn = str(tmp_path / "o.h5")
    h5out.save_all_data(fn, dict(beta=1.0), e, e * e, e, 1.0, 64, nf0=nf0, nfpi=nfpi)
    t = h5out.H5Reader(fn).tree()
    assert np.array_equal(t["/mc_data/energies"][:n_meas], e[:, 0]) and np.array_equal(t["/mc_data/energies"][n_meas:2 * n_meas], e[:, 1])
    assert np.array_equal(t["/mc_data/nf0"][:n_meas], nf0[:, 0])
    frep = stats.fstats_report(nf0, nfpi, stats.max_bin_depth(e.size))
    for name in ("nf_0", "nf_pi", "fsusc_0", "fsusc_pi", "binder_0", "binder_pi"):
        assert np.allclose(t["/stats/" + name], frep[name]["stats"]), name
    assert t["/binning/fsusc_pi"].shape[1] == 5 and "/binning/binder_0" not in t
    # DOS / IPR post-processing from the histories
    vol = 16
    sp = np.sort(rng.normal(size=(n_meas, n_chains, vol)), axis=2)
    ip = rng.uniform(0.2, 1.0, size=(n_meas, n_chains, vol))
    wg = np.linspace(-1, 1, 5)
    h5out.save_all_data(fn, dict(beta=1.0), e, e * e, e, 1.0, vol, spectrum_history=sp, ipr_history=ip, dos_wgrid=wg, dos_offset=0.05)
    t = h5out.H5Reader(fn).tree()
    assert t["/mc_data/spectrum_history"].shape == (vol, n_meas * n_chains) and np.array_equal(t["/mc_data/spectrum_history"][:, :n_meas], sp[:, 0].T)
    drep = stats.dos_report(sp, wg, 0.05, 1.0, stats.max_bin_depth(e.size))
    assert np.allclose(t["/stats/dos0"], drep["dos0"]["stats"]) and np.allclose(t["/stats/dos_err"], drep["dos_err"])
    assert t["/stats/ipr_err"].shape == (5, 3) and t["/stats/ipr0"].shape == (4,) and np.isfinite(t["/stats/nc"][1])


def test_cpp_data_save_header(tmp_path):
    """include/fk_mc_b200/data_save.hpp (C++ twin of stats.py + h5out.py): the file it writes is read back with the Python reader
    and its statistics are compared with the Python implementation on the same series."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe, fn = str(tmp_path / "data_save_test"), str(tmp_path / "cpp.h5")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "cpp", "data_save_test.cpp"),
                           "-o", exe])
    n, beta, vol = 512, 10.0, 64.0
    txt = subprocess.run([exe, fn, str(n), str(beta), str(vol)], check=True, capture_output=True, text=True).stdout
    # the same series in Python
    s, x = 88172645463325252, 0.0
    e, d2 = np.empty(n), np.empty(n)
    for i in range(n):
        s = (s * 6364136223846793005 + 1442695040888963407) % (1 << 64)
        u = (s >> 11) * (1.0 / 9007199254740992.0) - 0.5
        x = 0.7 * x + u
        e[i] = -0.25 + 0.01 * x
        d2[i] = 0.02 + 0.001 * u
    t = h5out.H5Reader(fn).tree()
    assert np.array_equal(t["/mc_data/energies"], e) and np.array_equal(t["/mc_data/d2energies"], d2)
    assert np.array_equal(t["/mc_data/ipr_history"], np.arange(12).reshape(3, 4) * 0.5)
    assert t["/parameters/beta"] == beta and t["/parameters/L"] == 8 and t["/parameters/output"] == "output.h5"
    rep = stats.energy_report(e, d2, beta, vol)
    for name, rtol in (("energy", 1e-12), ("d2energy", 1e-12), ("cv", 1e-6)):   # cv: a difference of nearly equal means, summed in another order
        assert np.allclose(t["/stats/" + name], rep[name]["stats"], rtol=rtol, atol=0)
        assert t["/binning/" + name].shape == (len(rep[name]["binning"]), 5)
        assert np.allclose(t["/binning/" + name][:, :4], np.array(rep[name]["binning"]), rtol=rtol, atol=0)
        assert np.allclose(t["/binning/" + name][:, 4], rep[name]["cor_length"], rtol=100 * rtol, atol=1e-9)
    printed = {ln.split()[0]: [float(v) for v in ln.split()[1:5]] for ln in txt.strip().splitlines()}
    assert np.allclose(printed["cv"], t["/stats/cv"]) and np.allclose(printed["c_energy"], t["/stats/c_energy"])
