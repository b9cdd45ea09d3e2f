"""
syncopy_b200 -- B200-native (sm_100a) engine for the per-trial spectral and
cross-spectral hot path of Syncopy's `freqanalysis` / `connectivityanalysis`.

    compute_functions   drop-in `computeFunction`s (per-trial, host arrays in/out)
    batched             whole-dataset entry points (device resident)
    engine              thin driver over the C ABI (libspyb200.so, include/spyb200.h)
    hostmath            per-call host-side constants (tapers, scales, frequency matching)

There is no CPU fallback: without the compiled library and an sm_100 GPU every
compute entry point raises `SpybError`.
"""
from ._lib import SpybError, LIB_PATH  # noqa: F401

__version__ = "0.1.0"
