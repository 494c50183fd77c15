"""B200-native Gauss-Newton solver for 2D bearing-only SLAM (hot path of
torchipeppo/prb-project-bearing-only-slam), behind a C ABI (include/bos_b200.h).

The package holds the CUDA kernels + C ABI (csrc/), a C++ host mirror of the reference's
framework API (host/), and this ctypes binding used by the tests and the benchmark.
Importing `capi.lib()` fails loudly when the CUDA library is not built; nothing here
computes on the CPU.
"""
from . import capi  # noqa: F401
from .problem import Problem, xyt_to_xycs  # noqa: F401
