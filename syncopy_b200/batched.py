"""
Whole-dataset ("batched") entry points: all trials of an `AnalogData`-like
array move through the GPU in one go instead of one `computeFunction` call per
trial.  This is the throughput path behind a `compute_sequential` override
(INTEGRATION.md); results are laid out exactly as the
reference's runtime would stack the per-trial results (trial k -> row block k).

Inputs are `[nTrials, nSamples, nChannels]` float32, either host `ndarray`s
(copied through pinned memory) or CUDA tensors.  Outputs are CUDA tensors
unless `to_host=True`.
"""
import os

import numpy as np
import torch

from . import hostmath as hm
from .engine import get_engine

# upper bound for the intermediate spectra buffer of the cross-spectral path
MAX_SPECTRA_BYTES = 24 << 30


def _device_trials(eng, trials):
    if isinstance(trials, torch.Tensor):
        assert trials.dim() == 3
        return trials.to(device=eng.tdev, dtype=torch.float32, non_blocking=True).contiguous()
    arr = np.ascontiguousarray(trials, dtype=np.float32)
    assert arr.ndim == 3, "trials must be [nTrials, nSamples, nChannels]"
    return torch.from_numpy(arr).to(eng.tdev, non_blocking=True)


def _freq_selection(nfft, samplerate, foi):
    freqs = np.fft.rfftfreq(nfft, 1 / samplerate)
    if foi is None:
        return freqs, None
    sel, idx = hm.best_match(freqs, foi, squash_duplicates=True)
    if idx.size == freqs.size and np.array_equal(idx, np.arange(freqs.size)):
        return freqs, None
    return sel, idx


def mtmfft(trials, samplerate, nSamples=None, taper="hann", taper_opt=None, demean_taper=False,
           ft_compat=False, foi=None, polyremoval=None, output="pow", keeptapers=True,
           keeptrials=True, to_host=False, engine=None):
    """
    All-trials version of `mtmfft_cF` (syncopy/specest/compRoutines.py:59-191).
    Returns (spec [nTrials | 1, nTaperOut, nFreq, nChannels], freqs).
    """
    eng = engine or get_engine()
    x = _device_trials(eng, trials)
    B, n_sig, _ = x.shape
    nfft = n_sig if nSamples is None else int(nSamples)
    freqs, fidx = _freq_selection(nfft, samplerate, foi)
    tapers = eng.taper_table(taper, n_sig, nfft, taper_opt)
    spec = eng.mtmfft(x, tapers, nfft, hm.mtmfft_scale(n_sig, nfft, ft_compat),
                      polyremoval=hm.polyremoval_code(polyremoval), demean_taper=demean_taper,
                      freq_idx=fidx, output=output, keeptapers=keeptapers)
    if not keeptrials:
        # runtime's trial mean: sum of the per-trial results / nTrials (computational_routine.py:1022-1032)
        spec = eng.sum_trials(spec, alpha=1.0 / B)[None]
    if to_host:
        spec = spec.cpu().numpy()
    return spec, freqs


def mtmconvol_frames(n_sig, nperseg, noverlap, boundary="zeros", padded=True):
    """(first frame start, number of frames kept) of `mtmconvol` for a trial of n_sig samples
    (syncopy/specest/mtmconvol.py:120-126,150; stft.py:101-127)."""
    hop = nperseg - noverlap
    n_keep = int(np.ceil(n_sig / hop))
    ext = n_sig
    if boundary is not None:
        frame_start0 = -(nperseg // 2)
        ext += 2 * (nperseg // 2)
    else:
        frame_start0 = 0
        n_keep -= nperseg
    if padded:
        ext += (-(ext - nperseg) % hop) % nperseg
    return frame_start0, max(0, min(n_keep, (ext - noverlap) // hop))


def mtmconvol(trials, samplerate, nperseg, noverlap, taper="hann", taper_opt=None, boundary="zeros",
              padded=True, foi=None, polyremoval=0, output="pow", keeptapers=True, to_host=False,
              engine=None):
    """
    All-trials version of the equidistant branch of `mtmconvol_cF`
    (syncopy/specest/compRoutines.py:386-390,410-413) for `soi = postselect = slice(None)`.
    Returns (spec [nTrials, nTime, nTaperOut, nFreq, nChannels], freqs).
    """
    eng = engine or get_engine()
    x = _device_trials(eng, trials)
    _, n_sig, _ = x.shape
    hop = nperseg - noverlap
    freqs, fidx = _freq_selection(nperseg, samplerate, foi)
    tapers = eng.taper_table(taper, nperseg, nperseg, taper_opt, periodic_dpss=True)
    frame_start0, n_frames = mtmconvol_frames(n_sig, nperseg, noverlap, boundary, padded)
    spec = eng.mtmconvol(x, tapers, nperseg, hop, frame_start0, n_frames, hm.stft_scale(nperseg),
                         polyremoval=hm.polyremoval_code(polyremoval), freq_idx=fidx, output=output,
                         keeptapers=keeptapers)
    if to_host:
        spec = spec.cpu().numpy()
    return spec, freqs


class CrossSpectraSum:
    """Trial-SUMMED cross spectra of one rank: `csd_sum / n_trials` is the trial average."""

    def __init__(self, csd_sum, n_trials, freqs):
        self.csd_sum, self.n_trials, self.freqs = csd_sum, n_trials, freqs

    def average(self, engine=None):
        eng = engine or get_engine()
        return eng.scale_(self.csd_sum.clone(), 1.0 / self.n_trials)[None]


def cross_spectra_sum(trials, samplerate=1, nSamples=None, foi=None, taper="hann", taper_opt=None,
                      demean_taper=False, polyremoval=False, engine=None, impl=0, out=None):
    """
    Sum over trials of the single-trial cross spectra of `cross_spectra_cF`
    (syncopy/connectivity/ST_compRoutines.py:268-424), i.e. what the runtime accumulates for
    `keeptrials=False` before dividing by nTrials.  Two kernels per trial chunk: the tapered FFT
    writes frequency-major spectra [nFreq][trial*taper][channel]; the contraction reduces them to
    [nFreq][channel][channel] in one pass over all (trial, taper) rows.
    """
    eng = engine or get_engine()
    x = _device_trials(eng, trials)
    B, n_sig, n_chan = x.shape
    nfft = n_sig if nSamples is None else int(nSamples)
    freqs, fidx = _freq_selection(nfft, samplerate, foi)
    n_freq = freqs.size
    tapers = eng.taper_table(taper, n_sig, nfft, taper_opt)
    K = tapers.shape[0]
    scale = hm.mtmfft_scale(n_sig, nfft)
    pr = hm.polyremoval_code(polyremoval)

    # impl: 0 auto (tcgen05 kernel when the shape is eligible), 1 CUDA-core FP32 kernel, 2 tcgen05 kernel
    use_tc = impl == 2 or (impl == 0 and eng.csd_planar_supported(n_chan))
    bytes_per_trial = n_freq * K * n_chan * 8
    chunk = max(1, min(B, MAX_SPECTRA_BYTES // max(1, bytes_per_trial)))
    acc = out
    rows = min(chunk, B) * K
    if use_tc:
        spectra = torch.empty((n_freq, rows, 2, n_chan), dtype=torch.float32, device=eng.tdev)
    else:
        spectra = torch.empty((n_freq, rows, n_chan), dtype=torch.complex64, device=eng.tdev)
    for b0 in range(0, B, chunk):
        nb = min(chunk, B - b0)
        view = spectra[:, :nb * K]
        beta = 0.0 if (b0 == 0 and out is None) else 1.0
        eng.mtmfft(x[b0:b0 + nb], tapers, nfft, scale, polyremoval=pr, demean_taper=demean_taper,
                   freq_idx=fidx, output="fourier_planar" if use_tc else "fourier", keeptapers=True,
                   out=view, freq_major=True)
        if use_tc:
            acc = eng.csd_accumulate_planar(view, acc=acc, alpha=1.0 / K, beta=beta)
        else:
            acc = eng.csd_accumulate(view, acc=acc, alpha=1.0 / K, beta=beta, impl=1)
    if acc is None:                        # no trials on this rank (more ranks than trials): an empty sum
        acc = torch.zeros((n_freq, n_chan, n_chan), dtype=torch.complex64, device=eng.tdev)
    return CrossSpectraSum(acc, B, freqs)


def cross_spectra(trials, samplerate=1, nSamples=None, foi=None, taper="hann", taper_opt=None,
                  demean_taper=False, polyremoval=False, keeptrials=False, to_host=False, engine=None,
                  impl=0):
    """
    `CrossSpectra` compute routine over all trials.  keeptrials=False returns the trial average
    [1, nFreq, C, C]; keeptrials=True stacks the single-trial results [nTrials, nFreq, C, C].
    """
    eng = engine or get_engine()
    if not keeptrials:
        res = cross_spectra_sum(trials, samplerate, nSamples, foi, taper, taper_opt, demean_taper,
                                polyremoval, engine=eng, impl=impl)
        csd, freqs = res.average(eng), res.freqs
    else:
        x = _device_trials(eng, trials)
        B, n_sig, n_chan = x.shape
        nfft = n_sig if nSamples is None else int(nSamples)
        freqs, fidx = _freq_selection(nfft, samplerate, foi)
        tapers = eng.taper_table(taper, n_sig, nfft, taper_opt)
        K = tapers.shape[0]
        csd = torch.empty((B, freqs.size, n_chan, n_chan), dtype=torch.complex64, device=eng.tdev)
        for b in range(B):
            spectra = eng.mtmfft(x[b:b + 1], tapers, nfft, hm.mtmfft_scale(n_sig, nfft),
                                 polyremoval=hm.polyremoval_code(polyremoval), demean_taper=demean_taper,
                                 freq_idx=fidx, output="fourier", keeptapers=True, freq_major=True)
            eng.csd_accumulate(spectra, acc=csd[b], alpha=1.0 / K, beta=0.0, impl=impl)
    if to_host:
        csd = csd.cpu().numpy()
    return csd, freqs


def _finish(result, to_host, out_host):
    """Optionally move a CUDA result to the host (into a caller-provided pinned tensor if given)."""
    if out_host is not None:
        out_host.copy_(result, non_blocking=True)
        torch.cuda.current_stream(result.device).synchronize()
        return out_host
    if to_host:
        return result.cpu().numpy()
    return result


def coherence(trials, samplerate=1, nSamples=None, foi=None, taper="hann", taper_opt=None,
              polyremoval=0, output="abs", to_host=False, engine=None, impl=0, reduce_group=None,
              out_host=None, gather=True):
    """
    `connectivityanalysis(method='coh')` compute chain: CrossSpectra(keeptrials=False) followed by
    NormalizeCrossSpectra (syncopy/connectivity/connectivity_analysis.py:460-473,549,587-599,677-679).
    The 1/nTrials of the trial mean is folded into the normalisation kernel.

    With `reduce_group` (a torch.distributed process group) `trials` is this rank's shard of the
    trial list: the trial-summed CSD and the trial count are all-reduced (the one collective of the
    path, replacing the reference's lock-serialised HDF5 `+=`, kwarg_decorators.py:722-735) and every
    rank returns the full result.  Returns (coh [1, nFreq, C, C], freqs).  With `gather=False` a rank keeps only
    the frequency slab it owns (coh [1, nFreq_slab, C, C] and that slab's frequencies) -- the layout in which each
    rank would write its part of the result file.
    """
    eng = engine or get_engine()
    n_chan = trials.shape[2]
    if impl in (0, 2) and eng.csd_planar_supported(n_chan) and not os.environ.get("SPYB_NO_TILES"):
        return _coherence_tiles(eng, trials, samplerate, nSamples, foi, taper, taper_opt, polyremoval, output,
                                to_host, reduce_group, out_host, gather)
    res = cross_spectra_sum(trials, samplerate, nSamples, foi, taper, taper_opt, False, polyremoval,
                            engine=eng, impl=impl)
    n_total = res.n_trials
    if reduce_group is not None:
        from .distributed import allreduce_csd
        n_total = allreduce_csd(res.csd_sum, res.n_trials, reduce_group, engine=eng)
    coh = eng.csd_normalize(res.csd_sum[None], output=output, pre_scale=1.0 / n_total)
    return _finish(coh, to_host, out_host), res.freqs


def _coherence_tiles(eng, trials, samplerate, nSamples, foi, taper, taper_opt, polyremoval, output, to_host,
                     reduce_group, out_host, gather):
    """
    128 / 256 channels: tapered FFT -> tcgen05 contraction that writes only the upper 128x128 tiles, each
    frequency straight into the slot buffer of the rank owning it (peer stores over NVLink for several ranks, see
    distributed.TileExchange) -> per-slab sum over ranks + normalisation.  The mirrored CSD is never materialised.
    """
    import torch.distributed as dist
    from .distributed import get_tile_exchange
    x = _device_trials(eng, trials)
    B, n_sig, n_chan = x.shape
    nfft = n_sig if nSamples is None else int(nSamples)
    freqs, fidx = _freq_selection(nfft, samplerate, foi)
    n_freq = freqs.size
    tapers = eng.taper_table(taper, n_sig, nfft, taper_opt)
    K = tapers.shape[0]
    scale = hm.mtmfft_scale(n_sig, nfft)
    pr = hm.polyremoval_code(polyremoval)
    chunk = max(1, min(B, MAX_SPECTRA_BYTES // max(1, n_freq * K * n_chan * 8)))
    spectra = eng.scratch("planar_spectra", (n_freq, min(chunk, B) * K, 2, n_chan), torch.float32)
    if reduce_group is None and chunk >= B and not os.environ.get("SPYB_NO_FUSED_COH"):
        # one rank, all rows in one launch: the contraction's epilogue normalises and mirrors (2 kernels in total)
        eng.mtmfft(x, tapers, nfft, scale, polyremoval=pr, freq_idx=fidx, output="fourier_planar",
                   keeptapers=True, out=spectra, freq_major=True)
        coh = eng.csd_coherence_planar(spectra, output=output)
        return _finish(coh[None], to_host, out_host), freqs
    ex = get_tile_exchange(eng, n_freq, n_chan, reduce_group)
    lo, hi = ex.f_begin[ex.rank], ex.f_begin[ex.rank + 1]
    if ex.world > 1 and B == 0:
        # more ranks than trials: this rank has nothing to contract, but its peers add the slots it owns in their
        # buffers and it still owns a frequency slab -- clear the slots, then sum the peers' tiles of the slab
        ex.clear_own_source()
        coh, _ = ex.finish(0, output=output)
        if not gather:
            return _finish(coh[None], to_host, out_host), freqs[lo:hi]
        return _finish(_gather_slabs(eng, ex, coh, reduce_group), to_host, out_host), freqs
    if ex.world > 1 and chunk >= B and not os.environ.get("SPYB_NO_FUSED_EXCHANGE"):
        # several ranks, all rows of the rank in one launch: the exchange is fused into the contraction on both
        # sides (peers' frequencies as tiles over NVLink, barrier, own slab with the peers' tiles added in the
        # normalising epilogue) -- no reduction or normalisation kernel
        eng.mtmfft(x, tapers, nfft, scale, polyremoval=pr, freq_idx=fidx, output="fourier_planar",
                   keeptapers=True, out=spectra, freq_major=True)
        ex.accumulate_others(spectra)
        coh, _ = ex.finish_fused(spectra, B, output=output)
        if not gather:
            return _finish(coh[None], to_host, out_host), freqs[lo:hi]
        return _finish(_gather_slabs(eng, ex, coh, reduce_group), to_host, out_host), freqs
    for b0 in range(0, B, chunk):
        nb = min(chunk, B - b0)
        view = spectra[:, :nb * K]
        eng.mtmfft(x[b0:b0 + nb], tapers, nfft, scale, polyremoval=pr, freq_idx=fidx, output="fourier_planar",
                   keeptapers=True, out=view, freq_major=True)
        # alpha = 1 on every rank (coherency does not depend on a common factor): a rank whose shard needs several
        # chunks stays compatible with peers that took the fused route above
        ex.accumulate(view, alpha=1.0, beta=0.0 if b0 == 0 else 1.0)
    if ex.world == 1 or not gather:
        out_dev = None
        coh, _ = ex.finish(B, output=output, out=out_dev)
        return _finish(coh[None], to_host, out_host), freqs[lo:hi]
    coh, _ = ex.finish(B, output=output)
    return _finish(_gather_slabs(eng, ex, coh, reduce_group), to_host, out_host), freqs


def _gather_slabs(eng, ex, coh, group):
    """Every rank wants the whole result: all-gather the normalised slabs (padded to the largest slab)."""
    import torch.distributed as dist
    lo, hi = ex.f_begin[ex.rank], ex.f_begin[ex.rank + 1]
    n_chan = ex.n_chan
    nf_max = max(ex.f_begin[r + 1] - ex.f_begin[r] for r in range(ex.world))
    pad = torch.zeros((nf_max, n_chan, n_chan), dtype=coh.dtype, device=eng.tdev)
    pad[:hi - lo] = coh
    full = torch.empty((ex.world, nf_max, n_chan, n_chan), dtype=coh.dtype, device=eng.tdev)
    if coh.dtype == torch.complex64:
        dist.all_gather_into_tensor(torch.view_as_real(full), torch.view_as_real(pad), group=group)
    else:
        dist.all_gather_into_tensor(full, pad, group=group)
    return torch.cat([full[r, :ex.f_begin[r + 1] - ex.f_begin[r]] for r in range(ex.world)], dim=0)[None]


def n_chan_of(csd):
    return int(csd.shape[-1])


def granger(trials, samplerate=1, nSamples=None, foi=None, taper="hann", taper_opt=None, polyremoval=0,
            rtol=5e-6, nIter=100, cond_max=1e4, to_host=False, engine=None, impl=0, reduce_group=None):
    """
    `connectivityanalysis(method='granger')` compute chain: CrossSpectra(keeptrials=False, demean_taper=True)
    followed by GrangerCausality (syncopy/connectivity/connectivity_analysis.py:576,864; AV_compRoutines.py:292-412).
    With `reduce_group`, `trials` is this rank's shard: the CSD sum is all-reduced, every rank regularises the
    average, the Wilson factorisation is sharded by frequency slab over the ranks (`spyb_wilson_sharded`) and the
    Granger slabs are gathered.  Returns (granger [1, nFreq, C, C] float32, metadata dict, freqs).
    """
    eng = engine or get_engine()
    res = cross_spectra_sum(trials, samplerate, nSamples, foi, taper, taper_opt, True, polyremoval,
                            engine=eng, impl=impl)
    n_total = res.n_trials
    if reduce_group is not None:
        from .distributed import allreduce_csd
        n_total = allreduce_csd(res.csd_sum, res.n_trials, reduce_group, engine=eng)
    csd_av = eng.scale_(res.csd_sum, 1.0 / n_total)
    reg, factor, ini_cn = eng.regularize_csd(csd_av, cond_max=cond_max, eps_max=1e-1)
    world = 1
    if reduce_group is not None:
        import torch.distributed as dist
        world = dist.get_world_size(reduce_group)
    if world > 1 and not os.environ.get("SPYB_WILSON_REPLICATED"):
        # every rank holds the same averaged CSD: factorise one frequency slab per rank (the plus operator is
        # replicated, its inputs are exchanged by broadcasts), then gather the Granger slabs
        import torch.distributed as dist
        from .distributed import WilsonExchange
        wx = WilsonExchange(eng, reg.shape[0], reduce_group)
        H, Sigma, conv, err, iters = eng.wilson_sf(reg, n_iter=nIter, rtol=rtol, slab=wx.slab, exchange=wx)
        lo, hi = wx.slab
        G = torch.zeros((reg.shape[0], n_chan_of(reg), n_chan_of(reg)), dtype=torch.float32, device=eng.tdev)
        if hi > lo:
            G[lo:hi] = eng.granger(reg[lo:hi].contiguous(), H[lo:hi].contiguous(), Sigma)
        for r in range(world):
            a, b = wx.f_begin[r], wx.f_begin[r + 1]
            if b > a:
                dist.broadcast(G[a:b], src=dist.get_global_rank(reduce_group, r), group=reduce_group)
        G = G[None]
    else:
        H, Sigma, conv, err, iters = eng.wilson_sf(reg, n_iter=nIter, rtol=rtol)
        G = eng.granger(reg, H, Sigma)[None]
    meta = {"converged--bool": np.array(conv), "max rel. err--float": np.array(err),
            "reg. factor--float": np.array(factor), "initial cond. num--float": np.array(np.float32(ini_cn)),
            "iterations": iters}
    return (G.cpu().numpy() if to_host else G), meta, res.freqs


def _conv_transform(eng, trials, plan_key, make_taps, polyremoval, output, to_host, trial_chunk):
    x = _device_trials(eng, trials)
    B, n_sig, n_chan = x.shape
    taps, expo = make_taps()
    plan = eng.conv_plan(plan_key, n_sig, taps, expo)
    kind_complex = hm.out_kind(output) == 2
    out = torch.empty((B, n_sig, 1, plan["n_scales"], n_chan),
                      dtype=torch.complex64 if kind_complex else torch.float32, device=eng.tdev)
    chunk = max(1, int(trial_chunk))
    for b0 in range(0, B, chunk):
        nb = min(chunk, B - b0)
        xd = eng.detrend(x[b0:b0 + nb], hm.polyremoval_code(polyremoval))
        eng.cwt(xd, plan, output=output, out=out[b0:b0 + nb].view(nb, n_sig, plan["n_scales"], n_chan))
    return out.cpu().numpy() if to_host else out


def wavelet(trials, samplerate, scales, wavelet, polyremoval=0, output="pow", to_host=False, engine=None,
            trial_chunk=32):
    """
    All-trials `wavelet_cF` (syncopy/specest/compRoutines.py:482-595) for `toi='all'`:
    [nTrials, nTime, 1, nScales, nChannels]; `wavelet` is any callable (t, s) -> psi (e.g. hostmath.Morlet()).
    """
    from .compute_functions import _plan_key
    eng = engine or get_engine()
    scales = np.asarray(scales, dtype=np.float64)
    dt = 1.0 / samplerate
    key = ("cwt", _plan_key(wavelet), scales.tobytes(), dt)
    return _conv_transform(eng, trials, key,
                           lambda: ([[hm.cwt_taps(wavelet, s, dt)] for s in scales], [[1.0] for _ in scales]),
                           polyremoval, output, to_host, trial_chunk)


def superlet(trials, samplerate, scales, order_max, order_min=1, c_1=3, adaptive=False, polyremoval=0,
             output="pow", to_host=False, engine=None, trial_chunk=32):
    """All-trials `superlet_cF` (syncopy/specest/compRoutines.py:654-762) for `toi='all'`."""
    from .compute_functions import superlet_tables
    eng = engine or get_engine()
    scales = np.asarray(scales, dtype=np.float64)
    dt = 1.0 / samplerate
    kw = dict(order_max=order_max, order_min=order_min, c_1=c_1, adaptive=adaptive)
    key = ("slt", scales.tobytes(), dt, tuple(sorted(kw.items())))
    return _conv_transform(eng, trials, key, lambda: superlet_tables(scales, dt, **kw), polyremoval, output,
                           to_host, trial_chunk)
