"""
TEST INFRASTRUCTURE (see oracle/__init__.py) -- CPU restatement of the wavelet
and superlet transforms of the reference:

    syncopy/specest/wavelets/wavelets.py:13-312     -> Morlet, Paul, DOG (time domain only)
    syncopy/specest/wavelets/transform.py:88-108    -> cwt_time
    syncopy/specest/wavelet.py:15-49                -> wavelet
    syncopy/specest/superlet.py:15-198,255-401      -> superlet (multiplicative + FASLT)
    syncopy/specest/compRoutines.py:482-595         -> wavelet_cF
    syncopy/specest/compRoutines.py:654-762         -> superlet_cF

Both transforms are "sample the mother wavelet on a finite support, then
scipy.signal.fftconvolve(data, kernel, mode='same') per scale", with the
complex128 result rounded to complex64; the superlet then forms complex64
geometric means across wavelet orders.
"""
import numpy as np
import scipy.special as _sp
from scipy.signal import fftconvolve

from .spectral import OUTPUT_DTYPES, convert_output, detrend_trial


# ---------------------------------------------------------------------------
# mother wavelets (time-domain forms only; `cwt_freq` is dead code upstream)
# ---------------------------------------------------------------------------

class Morlet:
    """wavelets.py:13-119: 'complete' Morlet, pi^-1/4 (e^{i w0 x} - e^{-w0^2/2}) e^{-x^2/2}."""

    def __init__(self, w0=6):
        self.w0 = w0

    def __call__(self, t, s=1.0):
        x = t / s
        return (np.exp(1j * self.w0 * x) - np.exp(-0.5 * self.w0 ** 2)) \
            * (np.exp(-0.5 * x ** 2) * np.pi ** (-0.25))

    def fourier_period(self, s):
        return 4 * np.pi * s / (self.w0 + (2 + self.w0 ** 2) ** 0.5)

    def scale_from_period(self, period):
        return period * (np.sqrt(self.w0 ** 2 + 2) + self.w0) / (4.0 * np.pi)


class Paul:
    """wavelets.py:122-219: 2^m i^m m! / sqrt(pi (2m)!) (1 - i x)^-(m+1)."""

    def __init__(self, m=4):
        self.m = m

    def __call__(self, t, s=1.0):
        m, x = self.m, t / s
        const = (2 ** m * 1j ** m * _sp.factorial(m)) / (np.pi * _sp.factorial(2 * m)) ** 0.5
        return const * (1 - 1j * x) ** -(m + 1)

    def fourier_period(self, s):
        return 4 * np.pi * s / (2 * self.m + 1)

    def scale_from_period(self, period):
        return period * (2 * self.m + 1) / (4 * np.pi)


class DOG:
    """wavelets.py:222-340: (-1)^(m+1)/sqrt(Gamma(m+1/2)) He_m(x) e^{-x^2/2} (real valued)."""

    def __init__(self, m=2):
        self.m = m

    def __call__(self, t, s=1.0):
        m, x = self.m, t / s
        const = (-1) ** (m + 1) / _sp.gamma(m + 0.5) ** 0.5
        return const * _sp.hermitenorm(m)(x) * np.exp(-x ** 2 / 2)

    def fourier_period(self, s):
        return 2 * np.pi * s / (self.m + 0.5) ** 0.5

    def scale_from_period(self, period):
        return period * np.sqrt(self.m + 0.5) / (2 * np.pi)


def support_times(M, dt):
    """`np.arange((-M + 1) / 2., (M + 1) / 2.) * dt` (transform.py:99, superlet.py:376-378)."""
    return np.arange((-M + 1) / 2.0, (M + 1) / 2.0) * dt


def cwt_kernel(wavelet, scale, dt):
    """Sampled, normalised CWT kernel for one scale (transform.py:96-103)."""
    t = support_times(10 * scale / dt, dt)
    return (dt ** 0.5 / (scale * 8 * np.pi)) * wavelet(t, scale)


def cwt(data, wavelet, scales, dt):
    """transform.py:88-108 with axis=0: complex64 [nScales, N, C]."""
    data = np.asarray(data)
    out = np.zeros((len(scales),) + data.shape, dtype=np.complex64)
    col = (slice(None),) + (None,) * (data.ndim - 1)
    for i, s in enumerate(scales):
        out[i] = fftconvolve(data, cwt_kernel(wavelet, s, dt)[col], mode="same")
    return out


def wavelet(data_arr, samplerate, scales, wavelet):
    """specest/wavelet.py:15-49."""
    return cwt(data_arr, wavelet, scales, 1 / samplerate)


def _select_time_post(spec, postselect):
    return spec.transpose(1, 0, 2)[postselect, :, :]


def wavelet_cF(trl_dat, preselect, postselect, toi=None, timeAxis=0, polyremoval=0,
               output="pow", noCompute=False, chunkShape=None, method_kwargs=None):
    """compRoutines.py:482-595."""
    dat = trl_dat.T if timeAxis != 0 else trl_dat
    n_time = toi.size if isinstance(toi, np.ndarray) else dat.shape[0]
    out_shape = (n_time, 1, method_kwargs["scales"].size, dat.shape[1])
    if noCompute:
        return out_shape, OUTPUT_DTYPES[output]
    dat = detrend_trial(np.array(dat), polyremoval)
    spec = wavelet(dat[preselect, :], **method_kwargs)
    spec = _select_time_post(spec, postselect)
    return convert_output(spec[:, None, :, :], output)


# ---------------------------------------------------------------------------
# superlets
# ---------------------------------------------------------------------------

class MorletSL:
    """superlet.py:255-299: B_c e^{i t/s} exp(-1/2 (k_sd t / (s 2 pi c))^2)."""

    def __init__(self, c_i=3, k_sd=5):
        self.c_i, self.k_sd = c_i, k_sd

    def __call__(self, t, s=1.0):
        ts = t / s
        B_c = self.k_sd / (s * self.c_i * (2 * np.pi) ** 1.5)
        return B_c * np.exp(1j * ts) * np.exp(-0.5 * (self.k_sd * ts / (2 * np.pi * self.c_i)) ** 2)


def superlet_kernel(c_i, scale, dt, k_sd=5):
    """Sampled kernel of `cwtSL` for one (cycle count, scale) (superlet.py:355-360,368-380)."""
    t = support_times(10 * scale * c_i / dt, dt)
    return (dt ** 0.5 / (4 * np.pi)) * MorletSL(c_i, k_sd)(t, scale)


def cwtSL(data, c_i, scales, dt):
    """superlet.py:321-365: complex64 [nScales, N, C]."""
    data = np.asarray(data)
    out = np.zeros((len(scales),) + data.shape, dtype=np.complex64)
    col = (slice(None),) + (None,) * (data.ndim - 1)
    for i, s in enumerate(scales):
        out[i] = fftconvolve(data, superlet_kernel(c_i, s, dt)[col], mode="same")
    return out


def compute_adaptive_order(freq, order_min, order_max):
    """superlet.py:383-401."""
    return order_min + (order_max - order_min) * (freq - freq[0]) / (freq[-1] - freq[0])


def multiplicative_slt(data, samplerate, scales, order_max, order_min=1, c_1=3):
    """superlet.py:108-126: prod_o cwtSL_o ** (1/nOrders), all in complex64."""
    dt = 1 / samplerate
    cycles = c_1 * np.arange(order_min, order_max + 1)
    n_ord = order_max + 1 - order_min
    gmean = np.power(cwtSL(data, cycles[0], scales, dt), 1 / n_ord)
    for c in cycles[1:]:
        gmean *= np.power(cwtSL(data, c, scales, dt), 1 / n_ord)
    return gmean


def faslt_plan(scales, order_max, order_min=1, c_1=3):
    """
    The per-scale bookkeeping of FASLT (superlet.py:139-174) as plain arrays:
    cycles of each wavelet set, exponents, alphas and the order jump positions.
    Shared with the GPU host code's *tests* only -- the product re-derives it.
    """
    fois = 1 / (2 * np.pi * scales)
    orders = compute_adaptive_order(fois, order_min, order_max)
    orders_int = np.int32(np.floor(orders))
    cycles = c_1 * np.unique(orders_int)
    exponents = 1 / (orders - order_min + 1)
    jumps = np.where(np.diff(orders_int))[0]
    alphas = orders % orders_int
    return cycles, exponents, alphas, jumps


def faslt(data, samplerate, scales, order_max, order_min=1, c_1=3):
    """superlet.py:129-198 (scales must be ordered high -> low, i.e. frequencies low -> high)."""
    dt = 1 / samplerate
    cycles, exponents, alphas, jumps = faslt_plan(scales, order_max, order_min, c_1)
    assert len(cycles) == len(jumps) + 1
    gmean = cwtSL(data, cycles[0], scales, dt)
    gmean = np.power(gmean.T, exponents).T
    last = 1
    for i, jump in enumerate(jumps):
        nxt = cwtSL(data, cycles[i + 1], scales[last:], dt)
        span = slice(last, jump + 1)
        gmean[span, :] *= np.power(nxt[: jump - last + 1].T, alphas[span] * exponents[span]).T
        gmean[jump + 1:] *= np.power(nxt[jump - last + 1:].T, exponents[jump + 1:]).T
        last = jump + 1
    return gmean


def superlet(data_arr, samplerate, scales, order_max, order_min=1, c_1=3, adaptive=False):
    """superlet.py:15-105."""
    if adaptive:
        return faslt(data_arr, samplerate, scales, order_max, order_min, c_1)
    return multiplicative_slt(data_arr, samplerate, scales, order_max, order_min, c_1)


def superlet_cF(trl_dat, preselect, postselect, toi=None, timeAxis=0, polyremoval=0,
                output="pow", noCompute=False, chunkShape=None, method_kwargs=None):
    """compRoutines.py:654-762."""
    dat = trl_dat.T if timeAxis != 0 else trl_dat
    n_time = toi.size if isinstance(toi, np.ndarray) else dat.shape[0]
    out_shape = (n_time, 1, method_kwargs["scales"].size, dat.shape[1])
    if noCompute:
        return out_shape, OUTPUT_DTYPES[output]
    dat = detrend_trial(np.array(dat), polyremoval)
    gmean = superlet(dat[preselect, :], **method_kwargs)
    gmean = _select_time_post(gmean, postselect)
    return convert_output(gmean[:, None, :, :], output)
