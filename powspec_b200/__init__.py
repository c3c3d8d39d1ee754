"""powspec_b200 — B200-native (sm_100a) replacement for powspec's hot path:
mass assignment -> r2c FFT -> window-deconvolved multipole binning.

The product is ``libpowspec_b200.so`` (hand-written CUDA + cuFFT behind a C ABI,
``include/powspec_b200.h``); this package is the thin Python host mirror used by
the tests and the benchmark.  There is no CPU fallback: without the compiled
library and a CUDA device every compute call raises.
"""
from .api import (  # noqa: F401
    Conf, Cata, Mesh, PK, PowspecB200Error, Context, genr_mesh, powspec, mesh_destroy,
    powspec_destroy, powspec_assign_names, load_library, library_path, run,
)
from .save_res import save_res  # noqa: F401
