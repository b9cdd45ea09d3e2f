"""
ctypes binding of libspyb200.so (the C ABI declared in include/spyb200.h).

There is no CPU fallback: if the shared library is missing or a call fails, a
`SpybError` is raised.  PyTorch tensors are only used as device buffers; the
library never sees a torch type.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libspyb200.so")


class SpybError(RuntimeError):
    """Raised when libspyb200 is unavailable or a kernel call reports an error."""


_lib = None

_vp, _i, _ll, _f, _d = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double
_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
# int exchange(void* ctx, int what, void* buf, long long row_bytes, int n_rows)  -- spyb_wilson_sharded
EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_int)

# name -> (restype, argtypes); must list every symbol of include/spyb200.h
PROTOTYPES = {
    "spyb_version": (_i, []),
    "spyb_init": (_i, [_i]),
    "spyb_last_error": (C.c_char_p, []),
    "spyb_launch_count": (_ll, []),
    "spyb_max_fft_len": (_i, [_i]),
    "spyb_mtmfft": (_i, [_vp, _i, _ll, _i, _i, _vp, _i, _i, _f, _i, _i, _vp, _i, _i, _i,
                         _vp, _ll, _ll, _ll, _vp, _vp]),
    "spyb_mtmconvol": (_i, [_vp, _i, _ll, _i, _i, _vp, _i, _i, _i, _i, _i, _f, _i, _vp, _i, _i, _i,
                            _vp, _ll, _ll, _ll, _ll, _vp]),
    "spyb_csd_accumulate": (_i, [_vp, _ll, _ll, _i, _i, _i, _vp, _i, _vp, _i, _f, _f, _vp, _i, _vp]),
    "spyb_csd_planar_supported": (_i, [_i, _ll, _ll]),
    "spyb_csd_accumulate_planar": (_i, [_vp, _ll, _ll, _i, _i, _i, _f, _f, _vp, _vp]),
    "spyb_csd_tile_count": (_i, [_i]),
    "spyb_csd_coherence_planar": (_i, [_vp, _ll, _ll, _i, _i, _i, _i, _vp, _vp]),
    "spyb_csd_accumulate_tiles": (_i, [_vp, _ll, _ll, _i, _i, _i, _f, _f, C.POINTER(C.c_void_p), _ip, _i, _i, _vp]),
    "spyb_csd_accumulate_tiles_others": (_i, [_vp, _ll, _ll, _i, _i, _i, _f, _f, C.POINTER(C.c_void_p), _ip, _i, _i, _vp]),
    "spyb_csd_coherence_planar_slots": (_i, [_vp, _ll, _ll, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp]),
    "spyb_csd_normalize_tiles": (_i, [_vp, _i, _i, _i, _f, _i, _vp, _vp]),
    "spyb_peer_alloc": (_i, [_ll, C.POINTER(C.c_void_p), C.c_char_p]),
    "spyb_peer_open": (_i, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "spyb_peer_memset": (_i, [_vp, _i, _ll, _vp]),
    "spyb_peer_close": (_i, [_vp]),
    "spyb_peer_free": (_i, [_vp]),
    "spyb_csd_normalize": (_i, [_vp, _ll, _i, _f, _i, _vp, _vp]),
    "spyb_detrend": (_i, [_vp, _i, _ll, _i, _i, _i, _vp, _ll, _vp]),
    "spyb_cwt": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "spyb_transpose": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "spyb_transpose_place": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _ll, _ll, _ll, _i, _vp]),
    "spyb_gather_rows": (_i, [_vp, _i, _ll, _vp, _i, _ll, _vp, _vp]),
    "spyb_scale": (_i, [_vp, _ll, _f, _vp]),
    "spyb_csd_mirror_upper": (_i, [_vp, _i, _i, _vp]),
    "spyb_sum_trials": (_i, [_vp, _i, _ll, _ll, _f, _f, _vp, _vp]),
    "spyb_axpby": (_i, [_vp, _vp, _f, _f, _vp, _ll, _vp]),
    "spyb_sqdev_accumulate": (_i, [_vp, _vp, _vp, _ll, _i, _vp]),
    "spyb_unit_accumulate": (_i, [_vp, _vp, _ll, _i, _vp]),
    "spyb_ppc_finish": (_i, [_vp, _vp, _ll, _i, _vp]),
    "spyb_xcov_kernel_spectra": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "spyb_xcov_finish": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "spyb_sosfilt": (_i, [_vp, _i, _ll, _i, _i, _dp, _i, _dp, _i, _i, _vp, _vp, _vp]),
    "spyb_upfirdn": (_i, [_vp, _i, _ll, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "spyb_standardize": (_i, [_vp, _i, _ll, _i, _i, _vp, _vp]),
    "spyb_rectify": (_i, [_vp, _vp, _ll, _vp]),
    "spyb_regularize_workspace_bytes": (_ll, [_i, _i]),
    "spyb_regularize_csd": (_i, [_vp, _i, _i, _d, _d, _i, _vp, _dp, _dp, _vp, _ll, _vp]),
    "spyb_wilson_workspace_bytes": (_ll, [_i, _i]),
    "spyb_wilson": (_i, [_vp, _i, _i, _i, _d, _vp, _vp, _ip, _dp, _ip, _vp, _ll, _vp]),
    "spyb_granger": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "spyb_wilson_sharded": (_i, [_vp, _i, _i, _i, _d, _vp, _vp, _ip, _dp, _ip, _vp, _ll, _i, _i, EXCHANGE_FN, _vp, _vp]),
}


def load(build_if_missing=False):
    """Load (once) and return the ctypes handle."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            from . import build as _build
            _build.build()
        else:
            raise SpybError(
                f"{LIB_PATH} not found -- run `python -m syncopy_b200.build` "
                "(there is no CPU fallback in this package)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise SpybError(f"libspyb200.so does not export `{name}`") from exc
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().spyb_last_error()
        raise SpybError(msg.decode("utf-8", "replace") if msg else f"libspyb200 error {rc}")


_initialised = set()


def init(device=0):
    """Select the device and verify it is an sm_100 part (raises otherwise)."""
    lib = load()
    if device not in _initialised:
        check(lib.spyb_init(int(device)))
        _initialised.add(device)
    return lib


def launch_count():
    return int(load().spyb_launch_count())
