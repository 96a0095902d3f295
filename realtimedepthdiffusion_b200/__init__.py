"""B200-native depth diffusion: Python host mirror over the C ABI (librtdd.so).

The product is the shared library (sm_100a kernels + `extern "C"` ABI + the
reference-named C++ functions).  This package is the thin host layer tests and
bench.py use: device memory and streams come from torch, every computation goes
through librtdd.so.  No CPU path exists here.
"""
from .api import (  # noqa: F401
    DepthDiffusion,
    RtddError,
    level_iterations,
    level_sizes,
    pitched_empty,
    pyramid_levels,
)
