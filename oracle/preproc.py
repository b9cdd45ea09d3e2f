"""
TEST INFRASTRUCTURE (see oracle/__init__.py) -- CPU restatement of the preprocessing compute functions
(SURVEY 8f row 4):

    syncopy/preproc/firws.py:13-222                      -> apply_fir, design_wsinc, windowed_sinc, invert_sinc, minphaserceps
    syncopy/preproc/resampling.py:14-138                 -> resample, downsample
    syncopy/preproc/compRoutines.py:27-148               -> sinc_filtering_cF
    syncopy/preproc/compRoutines.py:175-276              -> but_filtering_cF
    syncopy/preproc/compRoutines.py:303-338              -> rectify_cF
    syncopy/preproc/compRoutines.py:365-419              -> hilbert_cF
    syncopy/preproc/compRoutines.py:541-616              -> resample_cF
    syncopy/preproc/compRoutines.py:658-738              -> detrending_cF
    syncopy/preproc/compRoutines.py:765-832              -> standardize_cF

Third-party arithmetic (SciPy, like the reference): scipy.signal.convolve / butter / sosfilt / sosfiltfilt /
hilbert / resample_poly / detrend.  Pinned by tests/test_oracle_vs_reference.py against the reference's own files
(firws.py and resampling.py load by path; the cF bodies are compiled from compRoutines.py where it lies).
"""
import fractions

import numpy as np
import scipy.signal as sci
import scipy.signal.windows as sci_win

from .spectral import OUTPUT_DTYPES, convert_output


def apply_fir(data, fkernel, method="fft"):
    return sci.convolve(data, fkernel[:, None], mode="same", method=method)


def windowed_sinc(window, order, f_c):
    omega_c = 2 * np.pi * f_c
    win = getattr(sci_win, window)(order + 1)
    m_half = np.arange(1, order / 2 + 1)
    kernel = np.sin(omega_c * m_half) / m_half
    kernel = np.hstack([kernel[::-1], omega_c, kernel]) * win
    return kernel / kernel.sum()


def invert_sinc(kernel):
    kernel = -kernel
    kernel[len(kernel) // 2] += 1
    return kernel


def design_wsinc(window, order, f_c, filter_type="lp"):
    if order % 2 != 0:
        order += 1
    if filter_type == "lp":
        return windowed_sinc(window, order, f_c)
    elif filter_type == "hp":
        return invert_sinc(windowed_sinc(window, order, f_c))
    if filter_type == "bp":
        f_hp, f_lp = f_c
    elif filter_type == "bs":
        f_lp, f_hp = f_c
    lp_kernel = windowed_sinc(window, order, f_lp)
    hp_kernel = invert_sinc(windowed_sinc(window, order, f_hp))
    kernel = lp_kernel + hp_kernel
    if filter_type == "bp":
        kernel[len(kernel) // 2] -= 1
    return kernel


def minphaserceps(fkernel):
    n = len(fkernel)
    n_fft = int(2 ** np.ceil(np.log2(n * 1e3)))
    spec = np.abs(np.fft.fft(fkernel, n_fft))
    spec[spec < 1e-8] = 1e-8
    ceps = np.real(np.fft.ifft(np.log(spec)))
    ires = np.hstack([ceps[1:n_fft // 2], 0]) + np.conj(ceps[n_fft // 2:n_fft + 1][::-1])
    ceps = np.hstack([ceps[0], ires, np.zeros(n_fft // 2 - 2)])
    return np.real(np.fft.ifft(np.exp(np.fft.fft(ceps))))[:n]


def _detrend(dat, polyremoval):
    if polyremoval == 0:
        return sci.detrend(dat, type="constant", axis=0, overwrite_data=True)
    if polyremoval == 1:
        return sci.detrend(dat, type="linear", axis=0, overwrite_data=True)
    return dat


def sinc_filtering_cF(dat, samplerate=1, filter_type="lp", freq=None, order=None, window="hamming",
                      direction="onepass", polyremoval=None, timeAxis=0, noCompute=False, chunkShape=None):
    dat = dat.T if timeAxis != 0 else dat
    if noCompute:
        return dat.shape, np.float32
    dat = _detrend(dat, polyremoval)
    if order is None:
        order = dat.shape[0]
    fkernel = design_wsinc(window, order, freq / samplerate, filter_type)
    method = "direct" if np.any(np.isnan(dat)) else "fft"
    metadata = {"has_nan": np.array(method == "direct")}
    if direction == "onepass":
        filtered = apply_fir(dat, fkernel, method)
    elif direction == "twopass":
        filtered = apply_fir(apply_fir(dat, fkernel, method), fkernel, method)
    elif direction == "onepass-minphase":
        filtered = apply_fir(dat, minphaserceps(fkernel), method)
    return filtered, metadata


def but_filtering_cF(dat, samplerate=1, filter_type="lp", freq=None, order=6, direction="twopass", polyremoval=None,
                     timeAxis=0, noCompute=False, chunkShape=None):
    dat = dat.T if timeAxis != 0 else dat
    if noCompute:
        return dat.shape, np.float32
    metadata = {"has_nan": np.array(np.any(np.isnan(dat)))}
    dat = _detrend(dat, polyremoval)
    sos = sci.butter(order, freq, filter_type, fs=samplerate, output="sos")
    if direction == "twopass":
        return sci.sosfiltfilt(sos, dat, axis=0), metadata
    return sci.sosfilt(sos, dat, axis=0), metadata


def rectify_cF(dat, noCompute=False, chunkShape=None):
    if noCompute:
        return dat.shape, np.float32
    return np.abs(dat)


def hilbert_cF(dat, output="abs", timeAxis=0, noCompute=False, chunkShape=None):
    dat = dat.T if timeAxis != 0 else dat
    fmt = OUTPUT_DTYPES["fourier"] if output == "complex" else OUTPUT_DTYPES["abs"]
    if noCompute:
        return dat.shape, fmt
    return convert_output(sci.hilbert(dat, axis=0), output)


def _get_updn(orig_fs, new_fs):
    frac = fractions.Fraction.from_float(new_fs / orig_fs).limit_denominator()
    return frac.numerator, frac.denominator


def resample(data, orig_fs, new_fs, lpfreq=None, order=None):
    n = data.shape[0]
    up, down = _get_updn(orig_fs, new_fs)
    fs_ratio = new_fs / orig_fs
    if lpfreq is None:
        f_c = 0.5 * fs_ratio
    elif lpfreq == -1:
        f_c = None
    else:
        f_c = lpfreq / orig_fs
    if order is None:
        order = n * up
        order = 10000 if order > 10000 else order
    window = design_wsinc("hamming", order=order, f_c=f_c / up) if f_c else ("kaiser", 5.0)
    return sci.resample_poly(data, up, down, window=window, axis=0)


def resample_cF(dat, samplerate=1, new_samplerate=1, lpfreq=None, order=None, timeAxis=0, chunkShape=None,
                noCompute=False):
    dat = dat.T if timeAxis != 0 else dat
    if noCompute:
        return (int(np.ceil(dat.shape[0] * new_samplerate / samplerate)), dat.shape[1]), dat.dtype
    return resample(dat, samplerate, new_samplerate, lpfreq=lpfreq, order=order)


def detrending_cF(dat, polyremoval=None, timeAxis=0, noCompute=False, chunkShape=None):
    if polyremoval is None:
        return dat
    dat = dat.T if timeAxis != 0 else dat
    if noCompute:
        return dat.shape, np.float32
    return _detrend(dat, polyremoval), {"has_nan": np.array(np.any(np.isnan(dat)))}


def standardize_cF(dat, polyremoval=None, timeAxis=0, noCompute=False, chunkShape=None):
    dat = dat.T if timeAxis != 0 else dat
    if noCompute:
        return dat.shape, np.float32
    dat = _detrend(dat, polyremoval)
    return (dat - np.mean(dat, axis=0)) / np.std(dat, axis=0)
