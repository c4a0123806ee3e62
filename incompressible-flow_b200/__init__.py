"""impes-b200: B200-native IMPES hot path (pressure solve + saturation sub-steps).

Python here is test/bench plumbing over the C ABI (include/impes_b200.h); the product is
csrc/ (CUDA kernels + C ABI -> libimpes_b200.so) and host/ (C++ mirror of the reference's
PressureSolver / SaturationSolver / Param interfaces).  There is no CPU fallback: loading
fails loudly if the native library is missing.
"""
from . import abi  # noqa: F401
from . import binding, partition  # noqa: F401
from .binding import Solver, build_native, load_library, library_path  # noqa: F401
