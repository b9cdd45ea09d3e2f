"""
Drop-in preprocessing `computeFunction`s with the reference's names, signatures, dry-run
behaviour and return conventions (syncopy/preproc/compRoutines.py):

    sinc_filtering_cF  :27-148    windowed-sinc FIR filters (firws.py), 'same' convolution -> FFT convolution kernel
    but_filtering_cF   :175-276   Butterworth second-order sections, sosfilt / sosfiltfilt -> float64 recursion kernel
    rectify_cF         :303-338
    hilbert_cF         :365-419   analytic signal -> circular convolution with ifft(h) on the FFT convolution kernel
    downsample_cF      :446-500   (a strided view; no arithmetic)
    resample_cF        :541-616   resample_poly -> polyphase FIR kernel
    detrending_cF      :658-738
    standardize_cF     :765-832

Filter design (window functions, Butterworth poles, minimum-phase cepstrum, polyphase padding) is host-side float64
table generation (`hostmath`), exactly the reference's / SciPy's formulas; everything that touches the samples runs
in libspyb200.  NaNs: the reference switches its FIR filters to time-domain convolution when the trial contains
NaNs; an FFT convolution would smear them over the whole trial, so these functions raise instead of returning
something different from the reference (no CPU fallback).  Every cF has a batched sibling taking [B, N, C].
"""
import numpy as np
import torch

from . import _lib
from . import hostmath as hm
from .engine import get_engine


def _time_major(dat, timeAxis):
    return dat.T if timeAxis != 0 else dat


def _dev(eng, dat):
    return eng.to_device(np.ascontiguousarray(dat, dtype=np.float32))[None]


def _no_nans(dat, what):
    has_nan = bool(np.any(np.isnan(dat)))
    if has_nan:
        raise _lib.SpybError(f"{what}: the trial contains NaNs; the reference falls back to a time-domain "
                             f"convolution there, which this engine does not provide")
    return np.array(has_nan)


# ---- batched device-level versions ------------------------------------------------------------------------------------

def sinc_filter(x, samplerate, filter_type="lp", freq=None, order=None, window="hamming", direction="onepass",
                polyremoval=None, engine=None):
    """x [B, N, C] CUDA float32 -> filtered [B, N, C]."""
    eng = engine or get_engine()
    N = x.shape[1]
    if polyremoval in (0, 1):
        x = eng.detrend(x, int(polyremoval))
    if order is None:
        order = N
    fkernel = hm.design_wsinc(window, order, np.asarray(freq) / samplerate if np.ndim(freq) else freq / samplerate,
                              filter_type)
    if direction == "onepass-minphase":
        fkernel = hm.minphaserceps(fkernel)
    key = (window, int(order), filter_type, direction == "onepass-minphase", tuple(np.atleast_1d(freq / samplerate if
           np.ndim(freq) == 0 else np.asarray(freq) / samplerate).tolist()))
    y = eng.fir_same(x, fkernel, key)
    if direction == "twopass":
        y = eng.fir_same(y.contiguous(), fkernel, key)
    return y


def butterworth_filter(x, samplerate, filter_type="lp", freq=None, order=6, direction="twopass", polyremoval=None,
                       engine=None):
    eng = engine or get_engine()
    if polyremoval in (0, 1):
        x = eng.detrend(x, int(polyremoval))
    sos = hm.butter_sos(order, freq, filter_type, samplerate)
    return eng.sosfilt(x, sos, twopass=(direction == "twopass"))


def hilbert(x, output="abs", engine=None):
    eng = engine or get_engine()
    N = x.shape[1]
    out = "fourier" if output == "complex" else output
    return eng.fir_same(x, hm.analytic_kernel(N), ("hilbert", N), output=out)


def resample(x, samplerate, new_samplerate, lpfreq=None, order=None, engine=None):
    eng = engine or get_engine()
    h, up, down, m0, n_out = hm.resample_plan(x.shape[1], samplerate, new_samplerate, lpfreq, order)
    if up == down == 1:
        return x.clone()
    return eng.resample_poly(x, h, up, down, m0, n_out)


# ---- per-trial computeFunctions ---------------------------------------------------------------------------------------

def sinc_filtering_cF(dat, samplerate=1, filter_type="lp", freq=None, order=None, window="hamming",
                      direction="onepass", polyremoval=None, timeAxis=0, noCompute=False, chunkShape=None):
    dat = _time_major(dat, timeAxis)
    if noCompute:
        return dat.shape, np.float32
    metadata = {"has_nan": _no_nans(dat, "sinc_filtering_cF")}
    eng = get_engine()
    y = sinc_filter(_dev(eng, dat), samplerate, filter_type, freq, order, window, direction, polyremoval, eng)
    return y[0].cpu().numpy(), metadata


def but_filtering_cF(dat, samplerate=1, filter_type="lp", freq=None, order=6, direction="twopass", polyremoval=None,
                     timeAxis=0, noCompute=False, chunkShape=None):
    dat = _time_major(dat, timeAxis)
    if noCompute:
        return dat.shape, np.float32
    metadata = {"has_nan": np.array(np.any(np.isnan(dat)))}
    eng = get_engine()
    y = butterworth_filter(_dev(eng, dat), samplerate, filter_type, freq, order, direction, polyremoval, eng)
    return y[0].cpu().numpy(), metadata


def rectify_cF(dat, noCompute=False, chunkShape=None):
    if noCompute:
        return dat.shape, np.float32
    eng = get_engine()
    return eng.rectify(_dev(eng, dat))[0].cpu().numpy()


def hilbert_cF(dat, output="abs", timeAxis=0, noCompute=False, chunkShape=None):
    dat = _time_major(dat, timeAxis)
    fmt = np.complex64 if output == "complex" else np.float32
    if noCompute:
        return dat.shape, fmt
    _no_nans(dat, "hilbert_cF")
    eng = get_engine()
    return hilbert(_dev(eng, dat), output, eng)[0].contiguous().cpu().numpy()


def downsample_cF(dat, samplerate=1, new_samplerate=1, timeAxis=0, chunkShape=None, noCompute=False):
    dat = _time_major(dat, timeAxis)
    skipped = int(samplerate // new_samplerate)
    if noCompute:
        shape = list(dat.shape)
        shape[0] = int(np.ceil(dat.shape[0] / skipped))
        return tuple(shape), dat.dtype
    return dat[::skipped]                     # resampling.py:82-118: a strided view, no arithmetic to move


def resample_cF(dat, samplerate=1, new_samplerate=1, lpfreq=None, order=None, timeAxis=0, chunkShape=None,
                noCompute=False):
    dat = _time_major(dat, timeAxis)
    n = dat.shape[0]
    if noCompute:
        return (int(np.ceil(n * new_samplerate / samplerate)), dat.shape[1]), dat.dtype
    eng = get_engine()
    return resample(_dev(eng, dat), samplerate, new_samplerate, lpfreq, order, eng)[0].cpu().numpy()


def detrending_cF(dat, polyremoval=None, timeAxis=0, noCompute=False, chunkShape=None):
    if polyremoval is None:
        return dat
    dat = _time_major(dat, timeAxis)
    if noCompute:
        return dat.shape, np.float32
    has_nan = np.array(np.any(np.isnan(dat)))
    eng = get_engine()
    # NaN channels come back all-NaN from the sums, like the reference (:716-729)
    return eng.detrend(_dev(eng, dat), int(polyremoval))[0].cpu().numpy(), {"has_nan": has_nan}


def standardize_cF(dat, polyremoval=None, timeAxis=0, noCompute=False, chunkShape=None):
    dat = _time_major(dat, timeAxis)
    if noCompute:
        return dat.shape, np.float32
    eng = get_engine()
    x = _dev(eng, dat)
    if polyremoval in (0, 1):
        x = eng.detrend(x, int(polyremoval))
    return eng.standardize(x)[0].cpu().numpy()
