"""fk_mc_b200 -- B200-native (sm_100a) implementation of fk_mc's Metropolis weight-evaluation path.

The product is the C-ABI shared library ``fk_mc_b200/lib/libfkmc_b200.so`` (sources in
``fk_mc_b200/csrc``, interface in ``include/fkmc.h``).  This package is the thin Python binding
used by the tests and the benchmark; the C++ host mirror of the reference API lives in
``include/fk_mc_b200/``.  There is no CPU fallback: loading fails loudly when the library is
missing, and every compute call fails when no sm_100 GPU is present.
"""
from .binding import (  # noqa: F401
    ChainParams, Context, FkmcError, KINDS, LIB_PATH, build_library, cheb_sizes, exported_symbols, load_library, nccl_unique_id,
)

__all__ = ["ChainParams", "Context", "FkmcError", "KINDS", "LIB_PATH", "build_library", "cheb_sizes", "exported_symbols",
           "load_library", "nccl_unique_id"]
