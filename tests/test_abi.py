"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/fkmc.h
declares, and refuses to compute without a GPU (there is no CPU fallback)."""
import ctypes
import os
import re

import pytest

import fk_mc_b200 as fk
from conftest import HAVE_GPU, ROOT


def test_library_is_built_in_tree():
    assert os.path.exists(fk.LIB_PATH), "run python -c 'import __graft_entry__ as g; g.build()'"
    assert os.path.commonpath([fk.LIB_PATH, ROOT]) == ROOT


def test_exports_every_declared_symbol():
    lib = fk.load_library()
    syms = fk.exported_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "missing symbol " + s


def test_header_has_no_torch_or_cxx_types():
    text = open(os.path.join(ROOT, "include", "fkmc.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    assert "torch" not in code and "std::" not in code and "at::" not in code and "template" not in code
    assert 'extern "C"' in code
    # every entry point cites the reference interface it replaces
    assert len(re.findall(r"src/[a-z_/]+\.cpp:\d+", text)) >= 8


def test_product_never_touches_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "fk_mc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in text and "liboracle" not in text and "oracle/" not in text, f


@pytest.mark.skipif(HAVE_GPU, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    with pytest.raises(fk.FkmcError) as e:
        fk.Context("cubic2d", 8)
    assert e.value.code == 3  # FKMC_ERR_NO_DEVICE
    lib = fk.load_library()
    assert b"no CPU fallback" in lib.fkmc_last_error(None)


def test_null_context_is_rejected():
    lib = fk.load_library()
    assert lib.fkmc_volume(None) == -1
    assert lib.fkmc_chain_run_sweeps(None, 1) == 1  # FKMC_ERR_INVALID
    assert lib.fkmc_sync(None) == 1
    lib.fkmc_launch_count.restype = ctypes.c_int64
    assert lib.fkmc_launch_count(None) == -1


def test_cheb_sizes_host_logic():
    assert fk.cheb_sizes(64) == (10, 20) and fk.cheb_sizes(1024) == (16, 32) and fk.cheb_sizes(576, 2.5) == (16, 32)


def test_cpp_host_mirror_compiles_and_links(tmp_path):
    """include/fk_mc_b200/fk_mc.hpp (alps::params, fk::fk_mc<Lattice>, moves, measures) builds with plain g++ against the C ABI; on a machine
    without a GPU the driver program stops at lattice creation with the library's "no CPU fallback" error (no compute without a device)."""
    import subprocess
    exe = str(tmp_path / "host_api_test")
    libdir = os.path.join(ROOT, "fk_mc_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_api_test.cpp"), "-o", exe, "-L" + libdir, "-lfkmc_b200", "-Wl,-rpath," + libdir])
    if not HAVE_GPU:
        r = subprocess.run([exe, "8", "4", "4", "0", "0.5", "2", "32167", "0"], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stdout


def test_parameter_table_of_the_cpp_mirror(tmp_path):
    """fk_mc<L>::define_parameters carries the reference's names and defaults (include/fk_mc/fk_mc.hxx:177-207, src/mc_metropolis.cpp:11-19):
    checked on the host, no GPU involved."""
    import subprocess
    src = tmp_path / "params.cpp"
    src.write_text('''
#include <cstdio>
#include "fk_mc_b200/fk_mc.hpp"
int main() {
    fk::parameters_t p;
    p["seed"] = 42;                                   // a value set before define_parameters wins over the default
    fk::fk_mc<fk::hypercubic_lattice<2>>::define_parameters(p);
    printf("%ld %ld %ld %d\\n", long(p["nsweeps"]), long(p["sweep_len"]), long(p["ntherm_sweeps"]), int(bool(p["show_output"])));
    printf("%g %g %g %g\\n", double(p["beta"]), double(p["U"]), double(p["mu_c"]), double(p["mu_f"]));
    printf("%g %g %g %d %g\\n", double(p["mc_flip"]), double(p["mc_add_remove"]), double(p["mc_reshuffle"]), int(bool(p["cheb_moves"])), double(p["cheb_prefactor"]));
    printf("%d %ld %d %d %g %d\\n", int(bool(p["measure_history"])), long(p["Nf_start"]), int(bool(p["measure_ipr"])), int(bool(p["measure_eigenfunctions"])),
           double(p["cond_offset"]), int(bool(p["measure_stiffness"])));
    printf("%ld %ld %ld\\n", long(p["seed"]), long(p["SEED"]), long(p["nprocs"]));
    bool threw = false;
    try { const fk::parameters_t& q = p; (void)double(q["no_such_parameter"]); } catch (std::logic_error&) { threw = true; }
    printf("%d\\n", int(threw));
    return 0;
}
''')
    exe = str(tmp_path / "params")
    libdir = os.path.join(ROOT, "fk_mc_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), str(src), "-o", exe, "-L" + libdir, "-lfkmc_b200",
                           "-Wl,-rpath," + libdir])
    out = subprocess.check_output([exe]).decode().split("\n")
    assert out[0] == "1024 16 1 1"
    assert out[1] == "10 1 0.5 0.5"
    assert out[2] == "0 1 0 0 2.2"
    assert out[3] == "1 5 0 0 0.05 0"
    assert out[4] == "42 42 1"
    assert out[5] == "1"
