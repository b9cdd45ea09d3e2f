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


# ---------------------------------------------------------------------------
# wavelet / superlet kernels: sampled on the host in float64 exactly as the
# reference samples them, then handed to the GPU as convolution spectra
#   syncopy/specest/wavelets/wavelets.py:13-312   Morlet / Paul / DOG (time forms)
#   syncopy/specest/wavelets/transform.py:96-103  support + normalisation of cwt_time
#   syncopy/specest/superlet.py:255-299, 355-380  MorletSL, support + normalisation of cwtSL
#   syncopy/specest/superlet.py:108-198, 383-401  order / exponent bookkeeping
# ---------------------------------------------------------------------------

class Morlet:
    """Complete Morlet wavelet pi^-1/4 (exp(i w0 x) - exp(-w0^2/2)) exp(-x^2/2), x = t/s."""

    def __init__(self, w0=6):
        self.w0 = w0

    def __call__(self, t, s=1.0):
        x = t / s
        return (np.exp(1j * self.w0 * x) - np.exp(-0.5 * self.w0 ** 2)) * (np.pi ** -0.25 * np.exp(-0.5 * x ** 2))

    def fourier_period(self, s):
        return 4 * np.pi * s / (self.w0 + (2 + self.w0 ** 2) ** 0.5)

    def scale_from_period(self, period):
        return period * (np.sqrt(self.w0 ** 2 + 2) + self.w0) / (4.0 * np.pi)


class Paul:
    """Paul wavelet of order m."""

    def __init__(self, m=4):
        self.m = m

    def __call__(self, t, s=1.0):
        from scipy.special import factorial
        m, x = self.m, t / s
        const = (2 ** m * 1j ** m * factorial(m)) / (np.pi * factorial(2 * m)) ** 0.5
        return const * (1 - 1j * x) ** -(m + 1)

    def fourier_period(self, s):
        return 4 * np.pi * s / (2 * self.m + 1)

    def scale_from_period(self, period):
        return period * (2 * self.m + 1) / (4 * np.pi)


class DOG:
    """m-th derivative of a Gaussian (real valued)."""

    def __init__(self, m=2):
        self.m = m

    def __call__(self, t, s=1.0):
        from scipy.special import gamma, hermitenorm
        m, x = self.m, t / s
        return (-1) ** (m + 1) / gamma(m + 0.5) ** 0.5 * hermitenorm(m)(x) * np.exp(-x ** 2 / 2)

    def fourier_period(self, s):
        return 2 * np.pi * s / (self.m + 0.5) ** 0.5

    def scale_from_period(self, period):
        return period * np.sqrt(self.m + 0.5) / (2 * np.pi)


def support_times(M, dt):
    return np.arange((-M + 1) / 2.0, (M + 1) / 2.0) * dt


def cwt_taps(wavelet, scale, dt):
    """Sampled, normalised CWT kernel of one scale; `wavelet` is any callable (t, s) -> psi."""
    t = support_times(10 * scale / dt, dt)
    return (dt ** 0.5 / (scale * 8 * np.pi)) * np.asarray(wavelet(t, scale))


def superlet_taps(c_i, scale, dt, k_sd=5):
    """Sampled, normalised Morlet of `c_i` cycles as `cwtSL` builds it."""
    t = support_times(10 * scale * c_i / dt, dt)
    ts = t / scale
    B_c = k_sd / (scale * c_i * (2 * np.pi) ** 1.5)
    psi = B_c * np.exp(1j * ts) * np.exp(-0.5 * (k_sd * ts / (2 * np.pi * c_i)) ** 2)
    return (dt ** 0.5 / (4 * np.pi)) * psi


def superlet_factors(scales, order_max, order_min=1, c_1=3, adaptive=False):
    """
    Per scale the list of (cycle count, exponent) whose powers the superlet multiplies:
    multiplicative: every order with exponent 1/nOrders; FASLT: the fractional-order bookkeeping of the
    reference (first wavelet set on every scale, later sets from index `last` on, `alphas = orders % floor`).
    """
    scales = np.asarray(scales, dtype=np.float64)
    if not adaptive:
        cycles = c_1 * np.arange(order_min, order_max + 1)
        n_ord = order_max + 1 - order_min
        return [[(int(c), 1.0 / n_ord) for c in cycles] for _ in scales]
    fois = 1 / (2 * np.pi * scales)
    orders = order_min + (order_max - order_min) * (fois - fois[0]) / (fois[-1] - fois[0])
    orders_int = np.int32(np.floor(orders))
    cycles = c_1 * np.unique(orders_int)
    exponents = 1 / (orders - order_min + 1)
    jumps = np.where(np.diff(orders_int))[0]
    alphas = orders % orders_int
    if len(cycles) != len(jumps) + 1:
        raise ValueError("superlet: scales must be ordered high -> low for the adaptive transform")
    factors = [[(int(cycles[0]), float(exponents[i]))] for i in range(scales.size)]
    last = 1
    for k, jump in enumerate(jumps):
        for i in range(last, scales.size):
            a = alphas[i] * exponents[i] if i <= jump else exponents[i]
            factors[i].append((int(cycles[k + 1]), float(a)))
        last = jump + 1
    return factors


SMEM_FFT_MAX = 16384          # longest transform of the shared-memory FFT kernels (spyb_max_fft_len(1))


def conv_same_length(n_samples, taps_list):
    """Smallest supported circular length (power of two; 5-smooth beyond the shared-memory kernels) that reproduces
    fftconvolve(x, taps, 'same') for every kernel."""
    need = 16
    for taps in taps_list:
        M = len(taps)
        c0 = (M - 1) // 2
        need = max(need, n_samples + max(min(c0, n_samples - 1), min(M - 1 - c0, n_samples - 1)))
    L = 16
    while L < need:
        L <<= 1
    if L > SMEM_FFT_MAX:
        # beyond the shared-memory kernels the global-memory FFT takes any 5-smooth length: the smallest
        # 2^a 3^b 5^c >= need with a >= 4 (radix-16 passes first) instead of the next power of two
        best = L
        p5 = 1
        while p5 < best:
            p35 = p5
            while p35 < best:
                v = p35 * 16
                while v < need:
                    v <<= 1
                best = min(best, v)
                p35 *= 3
            p5 *= 5
        L = best
    return L


def conv_same_spectrum(taps, n_samples, L):
    """FFT_L(h) / L with h[d mod L] = taps[d + (M-1)//2], |d| < n_samples (taps that cannot reach an output are dropped)."""
    taps = np.asarray(taps, dtype=np.complex128)
    M = len(taps)
    c0 = (M - 1) // 2
    d = np.arange(-min(c0, n_samples - 1), min(M - 1 - c0, n_samples - 1) + 1)
    h = np.zeros(L, dtype=np.complex128)
    h[d % L] = taps[d + c0]
    return np.fft.fft(h) / L


def conv_segment_spectrum(taps, L, shift):
    """FFT_L(h) / L with h[d mod L] = taps[d + shift + (M-1)//2]: the 'same' convolution of a length-L segment,
    result index i' = i - shift (overlap-save blocks: the rows a launch keeps start at i' = 0)."""
    taps = np.asarray(taps, dtype=np.complex128)
    M = len(taps)
    assert M <= L
    j = np.arange(M)
    h = np.zeros(L, dtype=np.complex128)
    h[(j - shift - (M - 1) // 2) % L] = taps
    return np.fft.fft(h) / L


def conv_groups(n_samples, taps_per_scale, full_len, smem_max=SMEM_FFT_MAX, candidates=(4096, 8192)):
    """
    Split the scales of a 'same'-convolution transform into runs that share a circular length.  A kernel of M taps
    only needs its own support around every output sample, so short kernels run as overlap-save blocks of length Ls
    (each block yields V = Ls - (left extent) - (right extent) outputs) instead of one transform of the full padded
    length: n_seg * Ls log Ls operations instead of full_len log full_len, and smaller shared-memory tiles.
    Returns a list of dicts {s0, s1, L, seg, V, A, n_seg}: scales [s0, s1), `seg` False = one transform of full_len.
    Block lengths: measured on B200 at cfg-5 (profiles/cwt_segments_r2.log) 4096 / 8192 pay (superlets -20 %, wavelets
    unchanged: the transform kernel is not bound by its FFT passes), 1024 / 2048 do not (per-block overheads).
    """
    def cost(L, n):                                   # FFT passes + the load / multiply / store work of a transform
        c = n * L * (np.log2(L) + 6.0)
        return c * (3.0 if L > smem_max else 1.0)     # beyond the shared-memory kernels the transform lives in HBM

    if n_samples <= 0:
        return [dict(s0=0, s1=len(taps_per_scale), L=full_len, seg=False, V=n_samples, A=0, n_seg=1)]
    choice = []
    for tl in taps_per_scale:
        left = max(len(t) - 1 - (len(t) - 1) // 2 for t in tl)
        right = max((len(t) - 1) // 2 for t in tl)
        best, best_cost = 0, cost(full_len, 1)
        for Ls in candidates:
            V = Ls - left - right
            if Ls >= full_len or V < Ls // 4:
                continue
            c = cost(Ls, -(-n_samples // V))
            if c < 0.85 * best_cost:                  # not worth a separate launch group otherwise
                best, best_cost = Ls, c
        choice.append((best, left, right))
    groups, s0 = [], 0
    while s0 < len(choice):
        s1 = s0
        while s1 < len(choice) and choice[s1][0] == choice[s0][0]:
            s1 += 1
        Ls = choice[s0][0]
        if Ls == 0:
            groups.append(dict(s0=s0, s1=s1, L=full_len, seg=False, V=n_samples, A=0, n_seg=1))
        else:
            A = max(c[1] for c in choice[s0:s1])
            V = Ls - A - max(c[2] for c in choice[s0:s1])
            groups.append(dict(s0=s0, s1=s1, L=Ls, seg=True, V=V, A=A, n_seg=-(-n_samples // V)))
        s0 = s1
    return groups


# ---------------------------------------------------------------------------------------------------------
# preprocessing: host-side filter design (float64 tables, like the taper tables) for the kernels of csrc/preproc.cu
# ---------------------------------------------------------------------------------------------------------

def windowed_sinc(window, order, f_c):
    """syncopy/preproc/firws.py:109-148: windowed sinc low-pass of `order + 1` taps, unity gain."""
    import scipy.signal.windows as sci_win
    omega_c = 2 * np.pi * f_c
    win = getattr(sci_win, window)(order + 1)
    m_half = np.arange(1, order / 2 + 1)
    kernel = np.sin(omega_c * m_half) / m_half
    kernel = np.hstack([kernel[::-1], omega_c, kernel]) * win
    return kernel / kernel.sum()


def _invert_sinc(kernel):
    """firws.py:151-169: spectral inversion (low-pass -> high-pass); the kernel length is odd."""
    kernel = -kernel
    kernel[len(kernel) // 2] += 1
    return kernel


def design_wsinc(window, order, f_c, filter_type="lp"):
    """firws.py:49-106: low-, high-, band-pass and band-stop windowed-sinc kernels (f_c in units of the sampling rate)."""
    if order % 2 != 0:
        order += 1
    if filter_type == "lp":
        return windowed_sinc(window, order, f_c)
    if filter_type == "hp":
        return _invert_sinc(windowed_sinc(window, order, f_c))
    if filter_type == "bp":
        f_hp, f_lp = f_c
    elif filter_type == "bs":
        f_lp, f_hp = f_c
    else:
        raise ValueError(f"unknown filter type {filter_type!r}")
    kernel = windowed_sinc(window, order, f_lp) + _invert_sinc(windowed_sinc(window, order, f_hp))
    if filter_type == "bp":
        kernel[len(kernel) // 2] -= 1
    return kernel


def minphaserceps(fkernel):
    """firws.py:172-222: minimum-phase version of a FIR kernel through the real cepstrum (host, float64)."""
    n = len(fkernel)
    n_fft = int(2 ** np.ceil(np.log2(n * 1e3)))
    spec = np.abs(np.fft.fft(fkernel, n_fft))
    spec[spec < 1e-8] = 1e-8
    ceps = np.real(np.fft.ifft(np.log(spec)))
    ires = np.hstack([ceps[1:n_fft // 2], 0]) + np.conj(ceps[n_fft // 2:n_fft + 1][::-1])
    ceps = np.hstack([ceps[0], ires, np.zeros(n_fft // 2 - 2)])
    return np.real(np.fft.ifft(np.exp(np.fft.fft(ceps))))[:n]


def butter_sos(order, freq, filter_type, samplerate):
    """The design call of but_filtering_cF (compRoutines.py:263): second-order sections, float64 [S, 6]."""
    import scipy.signal as sci
    return np.ascontiguousarray(sci.butter(order, freq, filter_type, fs=samplerate, output="sos"), dtype=np.float64)


def sosfiltfilt_plan(sos):
    """Edge length and steady-state initial conditions exactly as scipy.signal.sosfiltfilt derives them
    (padtype='odd', padlen=None): ntaps = 2 S + 1 - min(#zero b2, #zero a2), padlen = 3 ntaps, zi = sosfilt_zi."""
    import scipy.signal as sci
    n_sections = sos.shape[0]
    ntaps = 2 * n_sections + 1
    ntaps -= min((sos[:, 2] == 0).sum(), (sos[:, 5] == 0).sum())
    return int(3 * ntaps), np.ascontiguousarray(sci.sosfilt_zi(sos), dtype=np.float64)


def analytic_kernel(n):
    """
    scipy.signal.hilbert multiplies the length-n spectrum by h (1 at DC / Nyquist, 2 at positive, 0 at negative
    frequencies): the analytic signal is the CIRCULAR convolution of x with a = ifft(h).  Returned: the 2n - 1 taps
    a[d mod n], d = -(n-1) .. n-1, so that a plain 'same' convolution (centre n - 1) reproduces the circular one.
    """
    h = np.zeros(n)
    if n % 2 == 0:
        h[0] = h[n // 2] = 1
        h[1:n // 2] = 2
    else:
        h[0] = 1
        h[1:(n + 1) // 2] = 2
    a = np.fft.ifft(h)
    d = np.arange(-(n - 1), n)
    return a[d % n]


def resample_plan(n_in, orig_fs, new_fs, lpfreq=None, order=None):
    """
    Filter and bookkeeping of `resample` (syncopy/preproc/resampling.py:14-79) + scipy.signal.resample_poly:
    returns (h float64 -- zero-padded and scaled by `up` as resample_poly does --, up, down, first kept row, n_out).
    """
    import fractions
    import math
    frac = fractions.Fraction.from_float(new_fs / orig_fs).limit_denominator()
    up, down = frac.numerator, frac.denominator
    fs_ratio = new_fs / orig_fs
    if lpfreq is None:
        f_c = 0.5 * fs_ratio
    elif lpfreq == -1:
        f_c = None
    else:
        f_c = lpfreq / orig_fs
    if order is None:
        order = n_in * up
        order = 10000 if order > 10000 else order
    if f_c:
        window = design_wsinc("hamming", order=order, f_c=f_c / up)
    else:
        window = None
    g = math.gcd(up, down)
    up //= g
    down //= g
    n_out = n_in * up
    n_out = n_out // down + bool(n_out % down)
    if window is not None:
        half_len = (window.size - 1) // 2
        h = np.array(window, dtype=np.float64)
    else:
        from scipy.signal import firwin
        max_rate = max(up, down)
        half_len = 10 * max_rate
        h = firwin(2 * half_len + 1, 1.0 / max_rate, window=("kaiser", 5.0)).astype(np.float32).astype(np.float64)
    h = h * up
    n_pre_pad = down - half_len % down
    n_post_pad = 0
    n_pre_remove = (half_len + n_pre_pad) // down

    def output_len(len_h):
        return (((n_in - 1) * up + len_h) - 1) // down + 1
    while output_len(h.size + n_pre_pad + n_post_pad) < n_out + n_pre_remove:
        n_post_pad += 1
    h = np.concatenate([np.zeros(n_pre_pad), h, np.zeros(n_post_pad)])
    return np.ascontiguousarray(h), int(up), int(down), int(n_pre_remove), int(n_out)
