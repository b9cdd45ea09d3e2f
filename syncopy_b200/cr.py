"""
Batched stand-in for `ComputationalRoutine.compute_sequential`
(syncopy/shared/computational_routine.py:944-1035) -- SURVEY 8(f) row 1.

The reference reads one trial from the source HDF5 dataset, calls the cF, and writes (or adds) the result into the
pre-allocated target dataset at `targetLayout[k]`, trial after trial.  Here the same contract is served with the
whole selection streamed through the GPU:

  * `source` is the dataset itself: any array-like `[nSamplesTotal, nChannels]` that supports NumPy-style row slicing
    (ndarray, `np.memmap`, an `h5py.Dataset` -- `AnalogData.data`), trials stacked along time;
  * `trialdefinition[t] = (start, stop, ...)` are the sample bounds of trial t (`AnalogData.sampleinfo`);
  * `trial_ids` is the selection-ordered trial list (`data.selection.trial_ids`): it may be permuted and may repeat
    trials; output block k belongs to `trial_ids[k]` and lands at rows `layout[k]` of the target exactly as
    `computational_routine.py:415-431,1017-1025` would place it (stacked along dim 0 in selection order);
  * `keeptrials=False` adds the per-trial results in selection order and divides by the trial count
    (`:1022-1032`);
  * trials of different lengths get one taper table per distinct length (windows are generated at the trial's own
    length, `mtmfft.py:99`) and must agree in output shape when averaged (`:318-320`).

Data movement: trials are packed into one of two pinned host buffers, copied to the device on a copy stream while the
previous chunk is being transformed, results return through two pinned buffers on a third stream and are written to
`target` while the next chunk computes.  Nothing here touches the oracle or any CPU fallback: the arithmetic is the
CUDA library's (`engine.Engine`).
"""
import numpy as np
import torch

from . import batched
from . import hostmath as hm
from .engine import get_engine

DEFAULT_CHUNK_BYTES = 512 << 20          # raw trial bytes per pipeline stage


class Layout:
    """Source / target geometry of one run (what `ComputationalRoutine.initialize` computes)."""

    def __init__(self, trialdefinition, trial_ids, n_chan):
        td = np.asarray(trialdefinition)
        self.bounds = [(int(td[t, 0]), int(td[t, 1])) for t in range(td.shape[0])]
        self.trial_ids = list(range(len(self.bounds))) if trial_ids is None else [int(t) for t in trial_ids]
        self.n_chan = int(n_chan)
        self.lengths = [self.bounds[t][1] - self.bounds[t][0] for t in self.trial_ids]
        self.source = [slice(*self.bounds[t]) for t in self.trial_ids]          # sourceLayout (time axis)

    def stack(self, rows_per_trial):
        """targetLayout: block k occupies rows [off_k, off_k + rows_k) of dim 0, in selection order."""
        out, off = [], 0
        for r in rows_per_trial:
            out.append(slice(off, off + int(r)))
            off += int(r)
        return out, off


def _groups_by_length(layout, chunk_bytes):
    """Chunks of selection positions with equal trial length, in selection order within a length."""
    by_len = {}
    for k, n in enumerate(layout.lengths):
        by_len.setdefault(n, []).append(k)
    chunks = []
    for n, ks in by_len.items():
        per = max(1, int(chunk_bytes // max(1, n * layout.n_chan * 4)))
        for i in range(0, len(ks), per):
            chunks.append((n, ks[i:i + per]))
    return chunks


_STAGING = {}          # (device, kind) -> list of two reusable buffers; pinning memory costs ~0.1 s per GB


def _staging(dev, kind, n_elems, dtype, pinned):
    """Two grow-only staging buffers per device and purpose, reused across calls."""
    key = (str(dev), kind, dtype)
    bufs = _STAGING.get(key)
    if bufs is None or bufs[0].numel() < n_elems:
        if pinned:
            bufs = [torch.empty(n_elems, dtype=dtype).pin_memory() for _ in range(2)]
        else:
            bufs = [torch.empty(n_elems, dtype=dtype, device=dev) for _ in range(2)]
        _STAGING[key] = bufs
    return bufs


def _pinned_view(source):
    """A torch view of `source` if it is host memory that CUDA can DMA from directly (pinned), else None."""
    try:
        if isinstance(source, torch.Tensor):
            t = source
        elif type(source) is np.ndarray and source.flags.writeable:     # (memmaps / read-only arrays are staged)
            t = torch.from_numpy(source)
        else:
            t = None
        if t is not None and t.dtype == torch.float32 and not t.is_cuda and t.is_contiguous() and t.is_pinned():
            return t
    except Exception:      # noqa: BLE001  (memmaps, read-only arrays, h5py datasets: staged through pinned buffers)
        pass
    return None


class _Pipeline:
    """Double-buffered pinned staging: H2D of chunk i+1 and D2H of chunk i-1 overlap the kernels of chunk i."""

    def __init__(self, eng, max_in_elems, max_out_elems, out_dtype, source=None):
        self.eng = eng
        self.dev = eng.tdev
        self.direct = _pinned_view(source)       # pinned source: DMA straight from it, no staging copy
        self.pin_in = None if self.direct is not None else _staging(self.dev, "in", max_in_elems, torch.float32, True)
        self.dev_in = _staging(self.dev, "dev", max_in_elems, torch.float32, False)
        self.pin_out = _staging(self.dev, "out", max_out_elems, out_dtype, True) if max_out_elems else None
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.back_stream = torch.cuda.Stream(self.dev)
        self.h2d_done = [None, None]
        self.in_free = [None, None]          # event: kernels that read dev_in[i] have finished
        self.d2h_done = [None, None]
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def upload(self, slot, source, layout, ks, n):
        """Pack trials ks (equal length n) into pinned buffer `slot`, start the H2D copy; returns the device view."""
        C = layout.n_chan
        nb = len(ks)
        dst = self.dev_in[slot][: nb * n * C].view(nb, n, C)
        if self.direct is not None:
            # pinned source: one DMA per run of selection entries that are adjacent in the dataset
            with torch.cuda.stream(self.copy_stream):
                if self.in_free[slot] is not None:
                    self.copy_stream.wait_event(self.in_free[slot])
                i = 0
                while i < nb:
                    j = i + 1
                    while j < nb and layout.source[ks[j]].start == layout.source[ks[j - 1]].stop:
                        j += 1
                    a, b = layout.source[ks[i]].start, layout.source[ks[j - 1]].stop
                    dst[i:j].view(-1, C).copy_(self.direct[a:b], non_blocking=True)
                    i = j
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            self.h2d_done[slot] = ev
            self.h2d_bytes += nb * n * C * 4
            return dst
        if self.h2d_done[slot] is not None:
            self.h2d_done[slot].synchronize()                # the pinned buffer is free again
        host = self.pin_in[slot][: nb * n * C].view(nb, n, C).numpy()
        for i, k in enumerate(ks):
            host[i] = source[layout.source[k], :]            # one read per (possibly repeated) selection entry
        with torch.cuda.stream(self.copy_stream):
            if self.in_free[slot] is not None:
                self.copy_stream.wait_event(self.in_free[slot])
            dst.copy_(self.pin_in[slot][: nb * n * C].view(nb, n, C), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.h2d_done[slot] = ev
        self.h2d_bytes += nb * n * C * 4
        return dst

    def wait_upload(self, slot):
        torch.cuda.current_stream(self.dev).wait_event(self.h2d_done[slot])

    def release_input(self, slot):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self.in_free[slot] = ev

    def download(self, slot, result):
        """Start the D2H copy of `result` (device) into pinned buffer `slot`; returns the host view (valid after
        `wait_download(slot)`)."""
        n = result.numel()
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(self.dev))
        host = self.pin_out[slot][:n].view(result.shape)
        with torch.cuda.stream(self.back_stream):
            self.back_stream.wait_event(done)
            host.copy_(result, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.back_stream)
        result.record_stream(self.back_stream)
        self.d2h_done[slot] = ev
        self.d2h_bytes += n * result.element_size()
        return host

    def wait_download(self, slot):
        if self.d2h_done[slot] is not None:
            self.d2h_done[slot].synchronize()


# ---- per-method chunk kernels: (device trials [B, n, C]) -> device result [B, rows, ...] -------------------------------

def _spec_fn(method, samplerate, cfg, eng):
    """Returns (fn(x_dev, n) -> [B, rows, ...] device tensor, rows(n), trailing shape(n), dtype, freq axis)."""
    output = cfg.get("output", "pow")
    cplx = hm.out_kind(output) == 2
    dt = torch.complex64 if cplx else torch.float32
    if method == "mtmfft":
        nS = cfg.get("nSamples")
        kw = dict(nSamples=nS, taper=cfg.get("taper", "hann"), taper_opt=cfg.get("taper_opt"),
                  demean_taper=cfg.get("demean_taper", False), ft_compat=cfg.get("ft_compat", False),
                  foi=cfg.get("foi"), polyremoval=cfg.get("polyremoval"), output=output,
                  keeptapers=cfg.get("keeptapers", True), engine=eng)

        def fn(x, n):
            spec, freqs = batched.mtmfft(x, samplerate, **kw)          # [B, K, nF, C]
            return spec, freqs
        return fn, (lambda n: 1), dt
    if method == "mtmconvol":
        kw = dict(taper=cfg.get("taper", "hann"), taper_opt=cfg.get("taper_opt"), boundary=cfg.get("boundary", "zeros"),
                  padded=cfg.get("padded", True), foi=cfg.get("foi"), polyremoval=cfg.get("polyremoval", 0),
                  output=output, keeptapers=cfg.get("keeptapers", True), engine=eng)

        def fn(x, n):
            spec, freqs = batched.mtmconvol(x, samplerate, cfg["nperseg"], cfg["noverlap"], **kw)   # [B, nT, K, nF, C]
            return spec, freqs
        return fn, None, dt
    if method == "wavelet":
        def fn(x, n):
            return batched.wavelet(x, samplerate, cfg["scales"], cfg["wavelet"], polyremoval=cfg.get("polyremoval", 0),
                                   output=output, engine=eng, trial_chunk=cfg.get("trial_chunk", 32)), None
        return fn, None, dt
    if method == "superlet":
        def fn(x, n):
            return batched.superlet(x, samplerate, cfg["scales"], cfg["order_max"], cfg.get("order_min", 1),
                                    cfg.get("c_1", 3), cfg.get("adaptive", False), polyremoval=cfg.get("polyremoval", 0),
                                    output=output, engine=eng, trial_chunk=cfg.get("trial_chunk", 32)), None
        return fn, None, dt
    raise ValueError(f"unknown method {method!r}")


def compute_sequential(source, trialdefinition, method, samplerate=1.0, trial_ids=None, keeptrials=True, target=None,
                       engine=None, chunk_bytes=DEFAULT_CHUNK_BYTES, **cfg):
    """
    Run `method` over the selected trials of `source` and place the results like the reference's runtime.

    method: 'mtmfft' | 'mtmconvol' | 'wavelet' | 'superlet'  (per-trial spectra, stacked along time) or
            'csd' | 'coh' | 'granger'                         (cross spectra; keeptrials=False chains)
    target: optional array-like to write into (ndarray / memmap / h5py dataset of the right shape and dtype);
            allocated as an ndarray when None.
    Returns dict(result=target, layout=[row slice per selection entry], freqs=..., meta=..., h2d_bytes, d2h_bytes).
    """
    eng = engine or get_engine()
    n_chan = int(source.shape[1])
    layout = Layout(trialdefinition, trial_ids, n_chan)
    if not layout.trial_ids:
        raise ValueError("empty trial selection")
    with torch.cuda.device(eng.device):
        if method in ("csd", "coh", "granger"):
            return _cross_spectral(eng, source, layout, method, samplerate, keeptrials, target, chunk_bytes, cfg)
        return _spectral(eng, source, layout, method, samplerate, keeptrials, target, chunk_bytes, cfg)


def _rows_per_trial(method, n, cfg):
    """Rows along the stacking (time) dimension of one trial's result -- the dry-run answer of the cF."""
    if method == "mtmfft":
        return 1
    if method == "mtmconvol":
        return batched.mtmconvol_frames(n, cfg["nperseg"], cfg["noverlap"], cfg.get("boundary", "zeros"),
                                        cfg.get("padded", True))[1]
    return n                                     # wavelet / superlet with toi='all'


def _spectral(eng, source, layout, method, samplerate, keeptrials, target, chunk_bytes, cfg):
    fn, _, dt = _spec_fn(method, samplerate, cfg, eng)
    np_dt = np.complex64 if dt == torch.complex64 else np.float32
    chunks = _groups_by_length(layout, chunk_bytes)
    max_in = max(n * len(ks) for n, ks in chunks) * layout.n_chan
    n_total = len(layout.trial_ids)
    rows = [_rows_per_trial(method, n, cfg) for n in layout.lengths]
    if not keeptrials and len(set(rows)) > 1:
        raise NotImplementedError("Averaging trials of unequal lengths in output currently not supported!")
    lay, total = layout.stack(rows)
    pipe = _Pipeline(eng, max_in, 0, dt, source)
    freqs, acc, trailing = None, None, None
    pending = None                                # (slot, ks, pinned host view): D2H in flight

    def land(p):
        slot, ks, host = p
        pipe.wait_download(slot)
        h = host.numpy()
        for i, k in enumerate(ks):
            target[lay[k]] = h[i]

    nxt = pipe.upload(0, source, layout, chunks[0][1], chunks[0][0])
    for ci, (n, ks) in enumerate(chunks):
        x, slot = nxt, ci % 2
        if ci + 1 < len(chunks):                  # host packing + H2D of the next chunk overlap the kernels below
            nxt = pipe.upload((ci + 1) % 2, source, layout, chunks[ci + 1][1], chunks[ci + 1][0])
        pipe.wait_upload(slot)
        spec, f = fn(x, n)
        pipe.release_input(slot)
        freqs = f if f is not None else freqs
        if method == "mtmfft":
            spec = spec[:, None]                  # [B, K, nF, C] -> one time row per trial
        if trailing is None:
            trailing = tuple(spec.shape[2:])
            if keeptrials:
                shape = (total,) + trailing
                if target is None:
                    target = np.empty(shape, dtype=np_dt)
                elif tuple(target.shape) != shape:
                    raise ValueError(f"target has shape {tuple(target.shape)}, the results need {shape}")
                per_chunk = max(len(k) * _rows_per_trial(method, m, cfg) for m, k in chunks) * int(np.prod(trailing))
                pipe.pin_out = _staging(pipe.dev, "out", per_chunk, dt, True)
        elif tuple(spec.shape[2:]) != trailing:
            raise ValueError("per-trial results disagree in their non-stacking dimensions")
        if not keeptrials:
            acc = eng.sum_trials(spec, acc=acc, alpha=1.0, beta=0.0 if acc is None else 1.0)
            continue
        if pending is not None and pending[0] == slot:
            land(pending)
            pending = None
        host = pipe.download(slot, spec)
        if pending is not None:
            land(pending)
        pending = (slot, ks, host)
    if pending is not None:
        land(pending)
    if not keeptrials:
        eng.scale_(acc, 1.0 / n_total)
        res = acc.cpu().numpy()
        pipe.d2h_bytes += res.nbytes
        if target is None:
            target = res
        else:
            target[...] = res
        lay = [slice(0, res.shape[0])] * n_total
    return dict(result=target, layout=lay, freqs=freqs, meta={}, h2d_bytes=pipe.h2d_bytes, d2h_bytes=pipe.d2h_bytes)


def _cross_spectral(eng, source, layout, method, samplerate, keeptrials, target, chunk_bytes, cfg):
    """
    'csd' : CrossSpectra (ST_compRoutines.py:268-424); keeptrials=True stacks [nTrials, nF, C, C]
    'coh' : CrossSpectra(keeptrials=False) + NormalizeCrossSpectra (connectivity_analysis.py:587-599,677-679)
    'granger': CrossSpectra(keeptrials=False, demean_taper=True) + GrangerCausality (:576,864)
    """
    n_chan = layout.n_chan
    taper, taper_opt = cfg.get("taper", "hann"), cfg.get("taper_opt")
    polyremoval = cfg.get("polyremoval", 0)
    nS = cfg.get("nSamples")
    demean = bool(cfg.get("demean_taper", method == "granger"))
    if nS is None and len(set(layout.lengths)) > 1:
        raise ValueError("trials of unequal length need a common padded length `nSamples` for cross spectra")
    nfft = int(nS) if nS is not None else layout.lengths[0]
    freqs, fidx = batched._freq_selection(nfft, samplerate, cfg.get("foi"))
    n_freq = freqs.size
    pr = hm.polyremoval_code(polyremoval)
    n_total = len(layout.trial_ids)
    if method == "csd" and keeptrials:
        # single-trial cross spectra: 8 nF C^2 bytes per trial, one trial per stage
        chunk_bytes = 1
    use_tc = eng.csd_planar_supported(n_chan)
    K = eng.taper_table(taper, layout.lengths[0], nfft, taper_opt).shape[0]
    per_trial_spec = n_freq * K * n_chan * 8
    chunk_bytes = min(chunk_bytes, max(1, (batched.MAX_SPECTRA_BYTES // max(1, per_trial_spec))) * layout.lengths[0] * n_chan * 4)
    chunks = _groups_by_length(layout, chunk_bytes)
    max_in = max(n * len(ks) for n, ks in chunks) * n_chan
    out_elems = n_freq * n_chan * n_chan if (method == "csd" and keeptrials) else 0
    pipe = _Pipeline(eng, max_in, out_elems, torch.complex64, source)
    max_rows = max(len(ks) for _, ks in chunks) * K
    # coherence on the tensor-core path with every row resident: the chunks' spectra fill one buffer and a single
    # contraction with normalising epilogue finishes the job (no cross-spectral matrix in memory)
    fused_coh = (method == "coh" and use_tc and n_total * per_trial_spec <= batched.MAX_SPECTRA_BYTES)
    if fused_coh:
        spectra = eng.scratch("cr_planar_spectra", (n_freq, n_total * K, 2, n_chan), torch.float32)
    elif use_tc:
        spectra = eng.scratch("cr_planar_spectra", (n_freq, max_rows, 2, n_chan), torch.float32)
    else:
        spectra = eng.scratch("cr_spectra", (n_freq, max_rows, n_chan), torch.complex64)
    acc = None
    lay, total = layout.stack([1] * n_total)
    if method == "csd" and keeptrials:
        shape = (total, n_freq, n_chan, n_chan)
        if target is None:
            target = np.empty(shape, dtype=np.complex64)
        elif tuple(target.shape) != shape:
            raise ValueError(f"target has shape {tuple(target.shape)}, the results need {shape}")
    pending = None
    row0 = 0
    nxt = pipe.upload(0, source, layout, chunks[0][1], chunks[0][0])
    for ci, (n, ks) in enumerate(chunks):
        x = nxt
        slot = ci % 2
        if ci + 1 < len(chunks):
            nxt = pipe.upload((ci + 1) % 2, source, layout, chunks[ci + 1][1], chunks[ci + 1][0])
        pipe.wait_upload(slot)
        tapers = eng.taper_table(taper, n, nfft, taper_opt)
        if fused_coh:
            view = spectra[:, row0: row0 + len(ks) * K]
            row0 += len(ks) * K
        else:
            view = spectra[:, : len(ks) * K]
        eng.mtmfft(x, tapers, nfft, hm.mtmfft_scale(n, nfft), polyremoval=pr, demean_taper=demean, freq_idx=fidx,
                   output="fourier_planar" if use_tc else "fourier", keeptapers=True, out=view, freq_major=True)
        pipe.release_input(slot)
        if fused_coh:
            continue
        if method == "csd" and keeptrials:
            one = (eng.csd_accumulate_planar(view, alpha=1.0 / K) if use_tc
                   else eng.csd_accumulate(view, alpha=1.0 / K, impl=1))
            if pending is not None:
                ps, pk, ph = pending
                pipe.wait_download(ps)
                target[lay[pk]] = ph.numpy()[None]
            pending = (slot, ks[0], pipe.download(slot, one))
            continue
        beta = 0.0 if acc is None else 1.0
        acc = (eng.csd_accumulate_planar(view, acc=acc, alpha=1.0 / K, beta=beta) if use_tc
               else eng.csd_accumulate(view, acc=acc, alpha=1.0 / K, beta=beta, impl=1))
    meta = {}
    if method == "csd" and keeptrials:
        if pending is not None:
            ps, pk, ph = pending
            pipe.wait_download(ps)
            target[lay[pk]] = ph.numpy()[None]
        return dict(result=target, layout=lay, freqs=freqs, meta=meta, h2d_bytes=pipe.h2d_bytes, d2h_bytes=pipe.d2h_bytes)
    if method == "csd":
        res = eng.scale_(acc, 1.0 / n_total)[None]
    elif fused_coh:
        res = eng.csd_coherence_planar(spectra, output=cfg.get("output", "abs"))[None]
    elif method == "coh":
        res = eng.csd_normalize(acc[None], output=cfg.get("output", "abs"), pre_scale=1.0 / n_total)
    else:
        csd_av = eng.scale_(acc, 1.0 / n_total)
        reg, factor, ini_cn = eng.regularize_csd(csd_av, cond_max=cfg.get("cond_max", 1e4), eps_max=1e-1)
        H, Sigma, conv, err, iters = eng.wilson_sf(reg, n_iter=cfg.get("nIter", 100), rtol=cfg.get("rtol", 5e-6))
        res = eng.granger(reg, H, Sigma)[None]
        meta = {"converged--bool": np.array(conv), "max rel. err--float": np.array(err),
                "reg. factor--float": np.array(factor), "initial cond. num--float": np.array(np.float32(ini_cn))}
    out_host = cfg.get("out_host")
    if out_host is not None:                       # caller-provided pinned tensor
        out_host.copy_(res, non_blocking=True)
        torch.cuda.current_stream(eng.tdev).synchronize()
        host = out_host
        nbytes = res.numel() * res.element_size()
    else:
        host = res.cpu().numpy()
        nbytes = host.nbytes
    if target is not None:
        target[...] = host if isinstance(host, np.ndarray) else host.numpy()
        host = target
    return dict(result=host, layout=[slice(0, 1)] * n_total, freqs=freqs, meta=meta, h2d_bytes=pipe.h2d_bytes,
                d2h_bytes=pipe.d2h_bytes + nbytes)
