"""
TEST INFRASTRUCTURE (see oracle/__init__.py) -- CPU restatement of the
(multi-)tapered FFT family of the reference:

    syncopy/specest/_norm_spec.py:9-46        -> taper_table, spectrum_scale
    syncopy/specest/mtmfft.py:16-129          -> mtmfft
    syncopy/specest/stft.py:16-159            -> stft
    syncopy/specest/mtmconvol.py:17-152       -> mtmconvol
    syncopy/shared/tools.py:224-343           -> best_match
    syncopy/shared/const_def.py:12-40         -> OUTPUT_DTYPES, convert_output
    syncopy/specest/compRoutines.py:59-191    -> mtmfft_cF
    syncopy/specest/compRoutines.py:244-414   -> mtmconvol_cF

The precision map of the reference is reproduced on purpose:
window x data and the FFT run in float64, the result is rounded to complex64
when stored, `mtmfft` scales *after* that rounding, `stft` scales *before* it,
whole-trial detrending stays in the input dtype.
"""
from hashlib import blake2b

import numpy as np
from scipy import signal as _sig

# syncopy/shared/const_def.py:12-22
OUTPUT_DTYPES = {
    "pow": np.float32, "abs": np.float32, "real": np.float32, "imag": np.float32,
    "angle": np.float32, "absreal": np.float32, "absimag": np.float32,
    "fourier": np.complex64, "complex": np.complex64,
}


def convert_output(spec, output):
    """syncopy/shared/const_def.py:25-40 (spectralConversions)."""
    if output == "pow":
        return (spec * np.conj(spec)).real.astype(np.float32)
    if output == "abs":
        return np.absolute(spec).real.astype(np.float32)
    if output in ("fourier", "complex"):
        return spec.astype(np.complex64)
    if output == "real":
        return np.real(spec).astype(np.float32)
    if output == "imag":
        return np.imag(spec).astype(np.float32)
    if output == "angle":
        return np.angle(spec).astype(np.float32)
    if output == "absreal":
        return np.abs(np.real(spec)).astype(np.float32)
    if output == "absimag":
        return np.abs(np.imag(spec)).astype(np.float32)
    raise ValueError(f"unknown output '{output}'")


def best_match(source, selection, squash_duplicates=False):
    """
    Nearest-element matching, restating the non-`span`, `tol=None` branch of
    syncopy/shared/tools.py:224-343 for a *sorted* `source` (Fourier frequency
    axes always are).  Ties go to the right neighbour (`left < right` is strict,
    tools.py:327), duplicates are squashed keeping first occurrences in query
    order (tools.py:333-336).
    """
    source = np.asarray(source)
    if np.issubdtype(type(selection), np.number):
        selection = [selection]
    selection = np.asarray(selection)
    pos = np.searchsorted(source, selection, side="left")
    left = np.abs(selection - source[np.maximum(pos - 1, 0)])
    right = np.abs(selection - source[np.minimum(pos, source.size - 1)])
    go_left = (pos == source.size) | (left < right)
    pos = pos.copy()
    pos[go_left] -= 1
    if squash_duplicates:
        _, first = np.unique(pos.astype(np.intp), return_index=True)
        pos = pos[np.sort(first)]
    return source[pos], pos


# ---------------------------------------------------------------------------
# tapers and normalisation
# ---------------------------------------------------------------------------

def taper_table(taper, length, n_padded, taper_opt=None):
    """
    float64 window table [K, length], scaled like `_norm_taper`
    (syncopy/specest/_norm_spec.py:27-46): dpss * sqrt(P), boxcar
    * sqrt(P / sum w), everything else * sqrt(4/3) * sqrt(P / sum w) where P is
    the *padded* length (mtmfft.py:99-101).  `taper=None` means boxcar
    (mtmfft.py:90-91).
    """
    if taper is None:
        taper = "boxcar"
    opt = dict(taper_opt or {})
    win = np.atleast_2d(getattr(_sig.windows, taper)(length, **opt)).astype(np.float64)
    if taper == "dpss":
        win = win * np.sqrt(n_padded)
    elif taper == "boxcar":
        win = win * np.sqrt(n_padded / win.sum())
    else:
        win = win * (np.sqrt(4 / 3) * np.sqrt(n_padded / win.sum()))
    return win


def spectrum_scale(n_norm):
    """`_norm_spec(..., mode='bins')`: sqrt(2) / n  (_norm_spec.py:9-24)."""
    return np.sqrt(2) / (n_norm * np.sqrt(1))


def mtmfft(data, samplerate, nSamples=None, taper="hann", taper_opt=None,
           demean_taper=False, ft_compat=False):
    """
    syncopy/specest/mtmfft.py:16-129.  data [N, C] (or [N]) -> complex64
    [K, nFreq, C], freqs.
    """
    data = np.asarray(data)
    if data.ndim < 2:
        data = data[:, None]
    n_sig = data.shape[0]
    n_pad = n_sig if nSamples is None else nSamples
    freqs = np.fft.rfftfreq(n_pad, 1 / samplerate)
    win = taper_table(taper, n_sig, n_pad, taper_opt)

    out = np.zeros((win.shape[0], freqs.size, data.shape[1]), dtype=np.complex64)
    norm_len = n_pad if ft_compat else n_sig * np.sqrt(n_pad / n_sig)
    for k in range(win.shape[0]):
        tapered = win[k][:, None] * data            # float64 product (mtmfft.py:112-113)
        if demean_taper:
            tapered = tapered - tapered.mean(axis=0)
        out[k] = np.fft.rfft(tapered, n=n_pad, axis=0)   # rounds to complex64 (:117)
        out[k] *= spectrum_scale(norm_len)               # scaled in complex64 (:119-127)
    return out, freqs


def detrend_trial(dat, polyremoval):
    """Whole-trial detrending as the cFs do it (compRoutines.py:169-172)."""
    # literal `==` tests as in the reference: `False == 0` de-means, `None` does nothing
    if polyremoval == 0:
        return _sig.detrend(dat, type="constant", axis=0)
    if polyremoval == 1:
        return _sig.detrend(dat, type="linear", axis=0)
    return dat


def freqs_hash(freqs):
    """compRoutines.py:181-183: blake2b hex digest of the frequency axis."""
    return np.array(blake2b(freqs).hexdigest().encode("utf-8"))


def mtmfft_cF(trl_dat, foi=None, timeAxis=0, keeptapers=True, polyremoval=None,
              output="pow", noCompute=False, chunkShape=None, method_kwargs=None):
    """syncopy/specest/compRoutines.py:59-191."""
    dat = trl_dat.T if timeAxis != 0 else trl_dat
    n_pad = method_kwargs["nSamples"]
    if n_pad is None:
        n_pad = dat.shape[0]
    freqs = np.fft.rfftfreq(n_pad, 1 / method_kwargs["samplerate"])
    _, fidx = best_match(freqs, foi, squash_duplicates=True)
    n_taper = method_kwargs["taper_opt"].get("Kmax", 1)
    out_shape = (1, max(1, n_taper * keeptapers), fidx.size, dat.shape[1])
    if noCompute:
        return out_shape, OUTPUT_DTYPES[output]

    dat = detrend_trial(np.array(dat), polyremoval)
    ftr, freqs = mtmfft(dat, **method_kwargs)
    spec = convert_output(ftr[None, :, fidx, :], output)
    meta = {"freqs_hash": freqs_hash(freqs)}
    if not keeptapers:
        return spec.mean(axis=1, keepdims=True), meta
    return spec, meta


# ---------------------------------------------------------------------------
# short-time FFT / mtmconvol
# ---------------------------------------------------------------------------

def stft(dat, fs=1.0, window=None, nperseg=256, noverlap=None, boundary="zeros",
         detrend=False, padded=True):
    """
    syncopy/specest/stft.py:16-159 for time-major input [N, C].
    Returns ftr [nFreq, C, nSeg] (complex128 -- the caller rounds), freqs.
    """
    x = np.moveaxis(np.asarray(dat), 0, -1)          # [C, N]
    if boundary is not None:                          # stft.py:101-105
        z = np.zeros(x.shape[:-1] + (nperseg // 2,), dtype=x.dtype)
        x = np.concatenate((z, x, z), axis=-1)
    if noverlap is None:
        noverlap = nperseg // 2
    hop = nperseg - noverlap
    if padded:                                        # stft.py:112-117 (float64 zeros!)
        nadd = (-(x.shape[-1] - nperseg) % hop) % nperseg
        x = np.concatenate((x, np.zeros(x.shape[:-1] + (nadd,))), axis=-1)
    nseg = (x.shape[-1] - noverlap) // hop
    starts = np.arange(nseg) * hop
    frames = x[..., starts[:, None] + np.arange(nperseg)[None, :]]   # [C, nSeg, nperseg]
    if detrend:
        frames = _sig.detrend(frames, type=detrend, axis=-1)
    if window is not None:
        frames = frames * window
    ftr = np.fft.rfft(frames, axis=-1)
    ftr = ftr * spectrum_scale(nperseg)               # before the c64 store (stft.py:154)
    return np.moveaxis(ftr, -1, 0), np.fft.rfftfreq(nperseg, 1 / fs)


def mtmconvol(data, samplerate, nperseg, noverlap=None, taper="hann", taper_opt=None,
              boundary="zeros", padded=True, detrend=False):
    """syncopy/specest/mtmconvol.py:17-152 -> complex64 [nTime, K, nFreq, C], freqs."""
    data = np.asarray(data)
    if data.ndim < 2:
        data = data[:, None]
    n = data.shape[0]
    if taper is None:
        taper = "boxcar"
    opt = dict(taper_opt or {})
    if taper == "dpss":
        opt["sym"] = False                            # mtmconvol.py:110-111
    win = taper_table(taper, nperseg, nperseg, opt)
    if noverlap is None:
        # `stft` defaults to half overlap; the nTime formula below needs a number
        noverlap = nperseg // 2
    hop = nperseg - noverlap
    n_time = int(np.ceil(n / hop))
    if boundary is None:
        n_time -= nperseg                             # mtmconvol.py:120-123
    freqs = np.fft.rfftfreq(nperseg, 1 / samplerate)
    out = np.zeros((n_time, win.shape[0], freqs.size, data.shape[1]), dtype=np.complex64)
    for k in range(win.shape[0]):
        pxx, _ = stft(data, samplerate, window=win[k], nperseg=nperseg, noverlap=noverlap,
                      boundary=boundary, padded=padded, detrend=detrend)
        out[:, k] = pxx.transpose(2, 0, 1)[:n_time]
    return out, freqs


def mtmconvol_cF(trl_dat, soi, postselect, equidistant=True, toi=None, foi=None, nTaper=1,
                 tapsmofrq=None, timeAxis=0, keeptapers=True, polyremoval=0, output="pow",
                 noCompute=False, chunkShape=None, method_kwargs=None):
    """syncopy/specest/compRoutines.py:244-414."""
    dat = trl_dat.T if timeAxis != 0 else trl_dat
    n_chan = dat.shape[1]
    if isinstance(toi, np.ndarray):
        n_time, bdry, pad = toi.size, None, False
    else:
        n_time = int(np.ceil(dat.shape[0] / (method_kwargs["nperseg"] - method_kwargs["noverlap"])))
        bdry, pad = "zeros", True
    taper_opt = method_kwargs["taper_opt"]
    if taper_opt:
        nTaper = taper_opt.get("Kmax", 1)
    out_shape = (n_time, max(1, nTaper * keeptapers), foi.size, n_chan)
    if noCompute:
        return out_shape, OUTPUT_DTYPES[output]

    if polyremoval == 0:          # compRoutines.py:376-381 (literal `==`: False de-means too)
        det = "constant"
    elif polyremoval == 1:
        det = "linear"
    else:
        det = False
    kw = dict(method_kwargs)
    kw.update(boundary=bdry, padded=pad, detrend=det)

    if equidistant:
        ftr, freqs = mtmconvol(np.asarray(dat)[soi, :], **kw)
        _, fidx = best_match(freqs, foi, squash_duplicates=True)
        spec = convert_output(ftr[postselect][:, :, fidx, :], output)
    else:
        spec = np.full((n_time, nTaper, foi.size, n_chan), np.nan, dtype=OUTPUT_DTYPES[output])
        for tk in range(len(soi)):
            ftr, freqs = mtmfft(np.asarray(dat)[soi[tk], :], kw["samplerate"],
                                taper=kw["taper"], taper_opt=taper_opt)
            _, fidx = best_match(freqs, foi, squash_duplicates=True)
            spec[tk] = convert_output(ftr[:, fidx, :], output)
    if not keeptapers:
        return np.nanmean(spec, axis=1, keepdims=True)
    return spec
