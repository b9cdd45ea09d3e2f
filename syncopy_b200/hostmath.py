"""
Host-side (NumPy/SciPy, float64) pieces of the hot path that are per-call, not
per-trial, work: window tables and their normalisation, frequency matching,
scaling constants, output dtype table.  The per-trial arithmetic lives in the
CUDA kernels.

Reference behaviour mirrored here:
    syncopy/specest/_norm_spec.py:9-46       taper / spectrum normalisation
    syncopy/specest/mtmfft.py:90-101         window construction for mtmfft
    syncopy/specest/mtmconvol.py:101-117     window construction for mtmconvol
    syncopy/shared/tools.py:224-343          best_match
    syncopy/shared/const_def.py:12-40        output kinds / dtypes
"""
import numpy as np
from scipy.signal import windows as _windows

# order = `out_kind` codes of include/spyb200.h
OUT_KINDS = {"pow": 0, "abs": 1, "fourier": 2, "complex": 2, "real": 3, "imag": 4,
             "angle": 5, "absreal": 6, "absimag": 7}

spectralDTypes = {
    "pow": np.float32, "abs": np.float32, "real": np.float32, "imag": np.float32,
    "angle": np.float32, "absreal": np.float32, "absimag": np.float32,
    "fourier": np.complex64, "complex": np.complex64,
}


def out_kind(output):
    try:
        return OUT_KINDS[output]
    except KeyError:
        raise ValueError(f"unsupported output '{output}'; expected one of {sorted(OUT_KINDS)}")


def normalized_tapers(taper, n_window, n_norm, taper_opt=None, force_periodic_dpss=False):
    """
    Window table [K, n_window] in float64, scaled as the reference does:
    `dpss` by sqrt(n_norm), `boxcar` by sqrt(n_norm / sum), any other window by
    sqrt(4/3) * sqrt(n_norm / sum).  For mtmfft `n_window` is the trial length
    and `n_norm` the padded length; for mtmconvol both are `nperseg` and dpss
    windows are generated with sym=False (`force_periodic_dpss`).
    """
    name = "boxcar" if taper is None else taper
    opts = dict(taper_opt) if taper_opt else {}
    if name == "dpss" and force_periodic_dpss:
        opts["sym"] = False
    table = np.atleast_2d(getattr(_windows, name)(n_window, **opts)).astype(np.float64)
    if name == "dpss":
        table = table * np.sqrt(n_norm)
    elif name == "boxcar":
        table = table * np.sqrt(n_norm / table.sum())
    else:
        table = table * (np.sqrt(4 / 3) * np.sqrt(n_norm / table.sum()))
    return table


def mtmfft_scale(n_signal, n_padded, ft_compat=False):
    """sqrt(2) / norm with norm = padded length (ft_compat) or L*sqrt(P/L) (mtmfft.py:119-127)."""
    norm = n_padded if ft_compat else n_signal * np.sqrt(n_padded / n_signal)
    return float(np.sqrt(2) / norm)


def stft_scale(nperseg):
    return float(np.sqrt(2) / nperseg)


def best_match(source, selection, squash_duplicates=False):
    """
    Closest-element lookup for sorted `source` (Fourier axes): returns
    (values, indices).  Ties resolve to the right-hand neighbour, an index past
    the end snaps to the last element; with `squash_duplicates` repeated hits
    are dropped keeping the first occurrence in query order.
    """
    src = np.asarray(source)
    if np.issubdtype(type(selection), np.number):
        selection = [selection]
    sel = np.asarray(selection)
    ins = np.searchsorted(src, sel, side="left")
    d_left = np.abs(sel - src[np.clip(ins - 1, 0, None)])
    d_right = np.abs(sel - src[np.clip(ins, None, src.size - 1)])
    idx = np.where((ins == src.size) | (d_left < d_right), ins - 1, ins)
    if squash_duplicates:
        _, first_pos = np.unique(idx.astype(np.intp), return_index=True)
        idx = idx[np.sort(first_pos)]
    return src[idx], idx


def polyremoval_code(polyremoval):
    """Map the cF's `polyremoval` (None / False / 0 / 1) onto the C ABI's -1 / 0 / 1 using the
    reference's literal `==` tests (False == 0 de-means)."""
    if polyremoval is None:
        return -1
    if polyremoval == 0:
        return 0
    if polyremoval == 1:
        return 1
    return -1
