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
