"""
Drop-in `computeFunction`s: same names, signatures, dry-run (`noCompute`)
behaviour and return conventions as the reference's per-trial middleware, with
the arithmetic executed by libspyb200 on the GPU.

    mtmfft_cF                   <- syncopy/specest/compRoutines.py:59-191
    mtmconvol_cF                <- syncopy/specest/compRoutines.py:244-414
    cross_spectra_cF            <- syncopy/connectivity/ST_compRoutines.py:268-424
    spectral_dyadic_product_cF  <- syncopy/connectivity/ST_compRoutines.py:29-117
    normalize_csd_cF            <- syncopy/connectivity/AV_compRoutines.py:35-112
    granger_cF                  <- syncopy/connectivity/AV_compRoutines.py:292-412
    cross_covariance_cF         <- syncopy/connectivity/ST_compRoutines.py:465-584
    ppc_column_cF               <- syncopy/connectivity/ST_compRoutines.py:158-233
    wavelet_cF                  <- syncopy/specest/compRoutines.py:482-595
    superlet_cF                 <- syncopy/specest/compRoutines.py:654-762

Each takes one trial as a host `ndarray` and returns a new host `ndarray`
(+ metadata dict where the reference returns one), so it can be bound with
`MultiTaperFFT.computeFunction = staticmethod(process_io(mtmfft_cF))` etc.
(INTEGRATION.md).  For throughput use `syncopy_b200.batched`, which keeps whole
datasets on the device.  Like the reference's cFs these functions do not
validate their inputs.
"""
from hashlib import blake2b

import numpy as np
import torch

from . import hostmath as hm
from .engine import get_engine


def _as_time_major(trl_dat, timeAxis):
    return trl_dat.T if timeAxis != 0 else trl_dat


def _freqs_hash(freqs):
    return np.array(blake2b(freqs).hexdigest().encode("utf-8"))


def _trial_to_device(eng, dat):
    """[N, C] host array (any float dtype / strides) -> [1, N, C] float32 CUDA tensor."""
    arr = np.ascontiguousarray(dat, dtype=np.float32)
    return eng.to_device(arr)[None]


# ---------------------------------------------------------------------------
# freqanalysis: mtmfft
# ---------------------------------------------------------------------------

def mtmfft_cF(trl_dat, foi=None, timeAxis=0, keeptapers=True, polyremoval=None, output="pow",
              noCompute=False, chunkShape=None, method_kwargs=None):
    dat = _as_time_major(trl_dat, timeAxis)
    n_sig, n_chan = dat.shape
    nfft = method_kwargs["nSamples"]
    if nfft is None:
        nfft = n_sig
    samplerate = method_kwargs["samplerate"]
    freqs = np.fft.rfftfreq(nfft, 1 / samplerate)
    _, freq_idx = hm.best_match(freqs, foi, squash_duplicates=True)
    taper_opt = method_kwargs.get("taper_opt") or {}
    n_taper = taper_opt.get("Kmax", 1)
    out_shape = (1, max(1, n_taper * keeptapers), freq_idx.size, n_chan)
    if noCompute:
        return out_shape, hm.spectralDTypes[output]

    eng = get_engine()
    x = _trial_to_device(eng, dat)
    tapers = eng.taper_table(method_kwargs.get("taper", "hann"), n_sig, nfft, taper_opt)
    scale = hm.mtmfft_scale(n_sig, nfft, method_kwargs.get("ft_compat", False))
    full = freq_idx.size == freqs.size and np.array_equal(freq_idx, np.arange(freqs.size))
    spec = eng.mtmfft(x, tapers, nfft, scale,
                      polyremoval=hm.polyremoval_code(polyremoval),
                      demean_taper=method_kwargs.get("demean_taper", False),
                      freq_idx=None if full else freq_idx,
                      output=output, keeptapers=keeptapers)
    res = spec.cpu().numpy().reshape(out_shape)
    return res, {"freqs_hash": _freqs_hash(freqs)}


# ---------------------------------------------------------------------------
# freqanalysis: mtmconvol
# ---------------------------------------------------------------------------

def _soi_as_range(soi, n):
    """`soi` is a slice (equidistant case); return (first sample, number of samples)."""
    start, stop, step = soi.indices(n)
    if step != 1:
        raise ValueError("mtmconvol_cF: strided sample selections are not supported")
    return start, max(0, stop - start)


def mtmconvol_cF(trl_dat, soi, postselect, equidistant=True, toi=None, foi=None, nTaper=1,
                 tapsmofrq=None, timeAxis=0, keeptapers=True, polyremoval=0, output="pow",
                 noCompute=False, chunkShape=None, method_kwargs=None):
    dat = _as_time_major(trl_dat, timeAxis)
    n_chan = dat.shape[1]
    nperseg, noverlap = method_kwargs["nperseg"], method_kwargs["noverlap"]
    hop = nperseg - noverlap
    if isinstance(toi, np.ndarray):
        n_time, boundary_zeros, padded = toi.size, False, False
    else:
        n_time = int(np.ceil(dat.shape[0] / hop))
        boundary_zeros, padded = True, True
    taper_opt = method_kwargs["taper_opt"]
    if taper_opt:
        nTaper = taper_opt.get("Kmax", 1)
    out_shape = (n_time, max(1, nTaper * keeptapers), foi.size, n_chan)
    if noCompute:
        return out_shape, hm.spectralDTypes[output]

    eng = get_engine()
    samplerate = method_kwargs["samplerate"]
    taper = method_kwargs["taper"]
    pr = hm.polyremoval_code(polyremoval)
    x = _trial_to_device(eng, dat)

    if equidistant:
        first, n_sel = _soi_as_range(soi, dat.shape[0])
        xs = x[:, first:first + n_sel, :]
        freqs = np.fft.rfftfreq(nperseg, 1 / samplerate)
        _, f_idx = hm.best_match(freqs, foi, squash_duplicates=True)
        tapers = eng.taper_table(taper, nperseg, nperseg, taper_opt, periodic_dpss=True)
        # number of frames the reference's stft produces and mtmconvol keeps (mtmconvol.py:120-150)
        n_keep = int(np.ceil(n_sel / hop))
        if boundary_zeros:
            frame_start0 = -(nperseg // 2)
        else:
            frame_start0 = 0
            n_keep -= nperseg
        ext = n_sel + (2 * (nperseg // 2) if boundary_zeros else 0)
        if padded:
            ext += (-(ext - nperseg) % hop) % nperseg
        n_seg = (ext - noverlap) // hop
        n_frames = max(0, min(n_keep, n_seg))
        # the taper mean (np.nanmean over the converted per-taper spectra, compRoutines.py:413) runs in the kernel:
        # K times fewer bytes come back (a NaN sample gives NaN either way)
        spec = eng.mtmconvol(xs, tapers, nperseg, hop, frame_start0, n_frames, hm.stft_scale(nperseg),
                             polyremoval=pr, freq_idx=f_idx, output=output, keeptapers=bool(keeptapers))
        spec = spec[0].cpu().numpy()                       # [nFrames, K or 1, nF, C]
        return spec[postselect]
    else:
        # one window per (non-equidistant) time point: a plain mtmfft of dat[soi[tk]] each
        # (compRoutines.py:392-408; note: no detrending and no padding in this branch)
        spec = np.full((n_time, nTaper, foi.size, n_chan), np.nan, dtype=hm.spectralDTypes[output])
        for tk, sl in enumerate(soi):
            first, n_sel = _soi_as_range(sl, dat.shape[0])
            freqs = np.fft.rfftfreq(n_sel, 1 / samplerate)
            _, f_idx = hm.best_match(freqs, foi, squash_duplicates=True)
            tapers = eng.taper_table(taper, n_sel, n_sel, taper_opt)
            res = eng.mtmfft(x[:, first:first + n_sel, :], tapers, n_sel, hm.mtmfft_scale(n_sel, n_sel),
                             freq_idx=f_idx, output=output, keeptapers=True)
            spec[tk] = res[0].cpu().numpy()
    if not keeptapers:
        return np.nanmean(spec, axis=1, keepdims=True)
    return spec


# ---------------------------------------------------------------------------
# connectivityanalysis: single-trial cross spectra
# ---------------------------------------------------------------------------

def cross_spectra_cF(trl_dat, samplerate=1, nSamples=None, foi=None, taper="hann", taper_opt=None,
                     demean_taper=False, polyremoval=False, timeAxis=0, chunkShape=None,
                     noCompute=False):
    dat = _as_time_major(trl_dat, timeAxis)
    n_sig, n_chan = dat.shape
    if nSamples is None:
        nSamples = n_sig
    freqs = np.fft.rfftfreq(nSamples, 1 / samplerate)
    if foi is not None:
        _, freq_idx = hm.best_match(freqs, foi, squash_duplicates=True)
        n_freq = freq_idx.size
    else:
        freq_idx, n_freq = None, freqs.size
    out_shape = (1, n_freq, n_chan, n_chan)
    if noCompute:
        return out_shape, hm.spectralDTypes["fourier"]

    eng = get_engine()
    x = _trial_to_device(eng, dat)
    tapers = eng.taper_table(taper, n_sig, nSamples, taper_opt)
    spectra = eng.mtmfft(x, tapers, nSamples, hm.mtmfft_scale(n_sig, nSamples),
                         polyremoval=hm.polyremoval_code(polyremoval), demean_taper=demean_taper,
                         freq_idx=freq_idx, output="fourier", keeptapers=True, freq_major=True)
    cs = eng.csd_accumulate(spectra, alpha=1.0 / tapers.shape[0])
    return cs.cpu().numpy()[None], {"freqs_hash": _freqs_hash(freqs)}


def spectral_dyadic_product_cF(specs, send_idx=None, send_N=None, rec_idx=None, rec_N=None,
                               chunkShape=None, noCompute=False):
    n_time, n_taper, n_freq, n_chan = specs.shape
    if send_idx is not None:
        out_shape = (n_time, n_freq, send_N, rec_N)
    else:
        out_shape = (n_time, n_freq, n_chan, n_chan)
    if noCompute:
        return out_shape, hm.spectralDTypes["fourier"]
    eng = get_engine()
    x = eng.to_device(np.ascontiguousarray(specs, dtype=np.complex64), dtype=torch.complex64)
    # [T, K, F, C] -> one "frequency" per (time, freq) pair with K rows each
    x = x.permute(0, 2, 1, 3).reshape(n_time * n_freq, n_taper, n_chan).contiguous()   # (a view when n_time == 1)
    if send_idx is not None:
        cs = eng.csd_accumulate(x, alpha=1.0 / n_taper, idx_i=np.asarray(send_idx), idx_j=np.asarray(rec_idx))
    else:
        cs = eng.csd_accumulate(x, alpha=1.0 / n_taper)
    return cs.cpu().numpy().reshape(out_shape)


def normalize_csd_cF(csd_av_dat, output="abs", chunkShape=None, noCompute=False):
    fmt = hm.spectralDTypes["fourier"] if output in ("complex", "fourier") else hm.spectralDTypes["abs"]
    if noCompute:
        return csd_av_dat.shape, fmt
    eng = get_engine()
    x = eng.to_device(np.ascontiguousarray(csd_av_dat, dtype=np.complex64), dtype=torch.complex64)
    return eng.csd_normalize(x, output=output).cpu().numpy()


# ---------------------------------------------------------------------------
# freqanalysis: wavelet / superlet
# ---------------------------------------------------------------------------

def granger_cF(csd_av_dat, rtol=5e-6, nIter=100, cond_max=1e4, chunkShape=None, noCompute=False):
    """
    Regularisation ladder -> Wilson factorisation -> Geweke-Granger causality of the trial-averaged cross
    spectra `csd_av_dat` [1, nFreq, C, C] (AV_compRoutines.py:378-412).  Returns float32 [1, nFreq, C, C] and the
    reference's metadata dict.
    """
    if noCompute:
        return csd_av_dat.shape, hm.spectralDTypes["abs"]
    eng = get_engine()
    csd = eng.to_device(np.ascontiguousarray(csd_av_dat[0], dtype=np.complex64), dtype=torch.complex64)
    reg, factor, ini_cn = eng.regularize_csd(csd, cond_max=cond_max, eps_max=1e-1)
    H, Sigma, conv, err, _ = eng.wilson_sf(reg, n_iter=nIter, rtol=rtol)
    G = eng.granger(reg, H, Sigma)
    meta = {
        "converged--bool": np.array(conv),
        "max rel. err--float": np.array(err),
        "reg. factor--float": np.array(factor),
        "initial cond. num--float": np.array(np.float32(ini_cn)),
    }
    return G[None].cpu().numpy(), meta


def cross_covariance_cF(trl_dat, samplerate=1, polyremoval=0, timeAxis=0, norm=False, fullOutput=False,
                        chunkShape=None, noCompute=False):
    """Single-trial cross-covariance / cross-correlation for all channel pairs (float32, the dry-run dtype)."""
    dat = _as_time_major(trl_dat, timeAxis)
    n, n_chan = dat.shape
    lags = np.arange(0, n // 2) if n % 2 == 0 else np.arange(0, n // 2 + 1)
    lags = lags * 1 / samplerate
    out_shape = (len(lags), 1, n_chan, n_chan)
    if noCompute:
        return out_shape, np.float32
    eng = get_engine()
    x = _trial_to_device(eng, dat)[0]
    cc = eng.cross_covariance(x, polyremoval=hm.polyremoval_code(polyremoval), norm=norm)
    res = cc.cpu().numpy()[:, None]
    return (res, lags) if fullOutput else res


def ppc_column_cF(cross_spectrum, trl2_idx=None, hdf5_path=None, chunkShape=None, noCompute=False, cross_spectrum2=None):
    """
    One trial pair of the PPC: cos(angle(z1 conj z2)).  The reference reads the second trial from the HDF5 file
    `hdf5_path` at `trl2_idx`; pass it as `cross_spectrum2` (or give a path -- opened with h5py if that is installed).
    The whole-dataset path is `syncopy_b200.statistics.ppc`, which needs no pair loop at all.
    """
    if noCompute:
        return cross_spectrum.shape, np.float32
    if cross_spectrum2 is None:
        import h5py                                   # only needed for the reference's file-based calling convention
        with h5py.File(hdf5_path, "r") as h5file:
            cross_spectrum2 = h5file["data"][trl2_idx]
    eng = get_engine()
    z1 = eng.to_device(np.ascontiguousarray(cross_spectrum, dtype=np.complex64), dtype=torch.complex64)
    z2 = eng.to_device(np.ascontiguousarray(cross_spectrum2, dtype=np.complex64), dtype=torch.complex64)
    acc = torch.empty_like(z1)
    eng.unit_accumulate(z1, acc, first=True)
    eng.unit_accumulate(z2, acc, first=False)
    # |u1 + u2|^2 = 2 + 2 cos(theta1 - theta2)
    return eng.ppc_finish(acc, 2).cpu().numpy()


def _plan_key(obj):
    """Hashable identity of a wavelet object (class name + public attributes)."""
    attrs = tuple(sorted((k, repr(v)) for k, v in vars(obj).items())) if hasattr(obj, "__dict__") else repr(obj)
    return (type(obj).__name__, attrs)


def _time_rows(sel, n):
    """Row indices selected by a slice / index array / list out of n rows, or None for 'all rows in order'."""
    if isinstance(sel, slice):
        start, stop, step = sel.indices(n)
        if (start, stop, step) == (0, n, 1):
            return None
        return np.arange(start, stop, step, dtype=np.int32)
    idx = np.asarray(sel)
    if idx.dtype == bool:
        idx = np.flatnonzero(idx)
    return np.where(idx < 0, idx + n, idx).astype(np.int32)


def _conv_transform_cF(trl_dat, preselect, postselect, toi, timeAxis, polyremoval, output, noCompute,
                       n_scales, plan_key, make_taps):
    dat = _as_time_major(trl_dat, timeAxis)
    n_time = toi.size if isinstance(toi, np.ndarray) else dat.shape[0]
    out_shape = (n_time, 1, n_scales, dat.shape[1])
    if noCompute:
        return out_shape, hm.spectralDTypes[output]
    eng = get_engine()
    x = eng.detrend(_trial_to_device(eng, dat), hm.polyremoval_code(polyremoval))
    rows_in = _time_rows(preselect, dat.shape[0])
    if rows_in is not None:
        if rows_in.size and not np.array_equal(rows_in, np.arange(rows_in[0], rows_in[0] + rows_in.size)):
            raise ValueError("wavelet / superlet: `preselect` must be a contiguous sample range")
        x = x[:, int(rows_in[0]):int(rows_in[0]) + rows_in.size, :] if rows_in.size else x[:, :0, :]
    n_sel = x.shape[1]
    taps, expo = make_taps()
    plan = eng.conv_plan(plan_key, n_sel, taps, expo)
    spec = eng.cwt(x, plan, output=output)                       # [1, n_sel, nScales, C]
    rows_out = _time_rows(postselect, n_sel)
    if rows_out is not None:
        spec = eng.gather_rows(spec, rows_out)
    return spec[0].cpu().numpy()[:, None, :, :]


def wavelet_cF(trl_dat, preselect, postselect, toi=None, timeAxis=0, polyremoval=0, output="pow",
               noCompute=False, chunkShape=None, method_kwargs=None):
    scales = np.asarray(method_kwargs["scales"], dtype=np.float64)
    wav = method_kwargs["wavelet"]
    dt = 1.0 / method_kwargs["samplerate"]

    def make_taps():
        return [[hm.cwt_taps(wav, s, dt)] for s in scales], [[1.0] for _ in scales]

    key = ("cwt", _plan_key(wav), scales.tobytes(), dt)
    return _conv_transform_cF(trl_dat, preselect, postselect, toi, timeAxis, polyremoval, output, noCompute,
                              scales.size, key, make_taps)


def superlet_cF(trl_dat, preselect, postselect, toi=None, timeAxis=0, polyremoval=0, output="pow",
                noCompute=False, chunkShape=None, method_kwargs=None):
    scales = np.asarray(method_kwargs["scales"], dtype=np.float64)
    dt = 1.0 / method_kwargs["samplerate"]
    kw = dict(order_max=method_kwargs["order_max"], order_min=method_kwargs.get("order_min", 1),
              c_1=method_kwargs.get("c_1", 3), adaptive=method_kwargs.get("adaptive", False))

    def make_taps():
        return superlet_tables(scales, dt, **kw)

    key = ("slt", scales.tobytes(), dt, tuple(sorted(kw.items())))
    return _conv_transform_cF(trl_dat, preselect, postselect, toi, timeAxis, polyremoval, output, noCompute,
                              scales.size, key, make_taps)


def superlet_tables(scales, dt, order_max, order_min=1, c_1=3, adaptive=False):
    """(taps[s][j], exponents[s][j]) of a superlet transform; zero exponents (z**0 = 1) are dropped."""
    factors = hm.superlet_factors(scales, order_max, order_min, c_1, adaptive)
    taps, expo = [], []
    for s, fl in zip(scales, factors):
        fl = [(c, a) for (c, a) in fl if a != 0.0] or [fl[0]]
        taps.append([hm.superlet_taps(c, s, dt) for c, _ in fl])
        expo.append([a for _, a in fl])
    return taps, expo
