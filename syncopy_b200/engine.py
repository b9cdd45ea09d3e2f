"""
Device-side driver of the per-trial spectral engine.

`Engine` owns nothing but small cached tables (tapers, frequency index lists) on
one GPU; trial data and results live in caller-provided or freshly allocated
PyTorch CUDA tensors that are handed to libspyb200 as raw device pointers.
Everything is enqueued on torch's current stream of the device.
"""
import functools
import os
import threading
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from . import hostmath as hm

# bounds of the per-engine device-side caches (entries): taper tables / index lists, convolution plans
MAX_TABLES = 64
MAX_PLANS = 8
MAX_LAUNCH_DIM = 65535                      # CUDA grid.y / grid.z limit of one launch
MAX_LONG_FFT = 1 << 24                      # longest transform of the global-memory FFT path (csrc/fft_long.cu)

_CDTYPE = {True: torch.complex64, False: torch.float32}


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _trial_layout(x):
    """Validate a [B, N, C] float32 CUDA stack of dense time-major trials; return the trial stride."""
    assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 3
    B, N, Cn = x.shape
    # (strides of size-1 dimensions are meaningless in torch, so test the 2-D slices instead)
    assert B == 0 or x[0].is_contiguous(), "trials must be dense [sample][channel]"
    return x.stride(0) if B > 1 else N * Cn


def _on_own_device(cls):
    """Run every public method with the engine's device current: the library launches on the current device."""
    def wrap(fn):
        @functools.wraps(fn)
        def inner(self, *a, **kw):
            if torch.cuda.current_device() == self.device:
                return fn(self, *a, **kw)
            with torch.cuda.device(self.device):
                return fn(self, *a, **kw)
        return inner
    for name, attr in list(vars(cls).items()):
        if callable(attr) and not name.startswith("_"):
            setattr(cls, name, wrap(attr))
    return cls


@_on_own_device
class Engine:
    def __init__(self, device=0):
        if not torch.cuda.is_available():
            raise _lib.SpybError("no CUDA device visible -- syncopy_b200 has no CPU fallback")
        self.device = int(device)
        self.lib = _lib.init(self.device)
        self.tdev = torch.device("cuda", self.device)
        self._tables = {}                 # grow-only scratch buffers (few, reused)
        self._lru = OrderedDict()         # taper tables, index lists: bounded, least recently used out first
        self._plans = OrderedDict()       # convolution plans (tens of MB each): bounded
        self._lock = threading.RLock()    # the process-wide engine is shared by worker threads

    # ------------------------------------------------------------------ utilities
    def stream(self):
        return torch.cuda.current_stream(self.tdev).cuda_stream

    def synchronize(self):
        torch.cuda.synchronize(self.tdev)

    def to_device(self, arr, dtype=torch.float32):
        """Host ndarray / torch tensor -> contiguous CUDA tensor of `dtype` on this device."""
        if isinstance(arr, torch.Tensor):
            return arr.to(device=self.tdev, dtype=dtype).contiguous()
        a = np.ascontiguousarray(arr)
        return torch.from_numpy(a).to(device=self.tdev, dtype=dtype, non_blocking=False).contiguous()

    def _cached(self, cache, limit, key, make):
        with self._lock:
            if key in cache:
                cache.move_to_end(key)
                return cache[key]
        val = make()
        with self._lock:
            cache[key] = val
            cache.move_to_end(key)
            while len(cache) > limit:
                cache.popitem(last=False)     # tensors still referenced by a caller stay alive until released
        return val

    def taper_table(self, taper, n_window, n_norm, taper_opt=None, periodic_dpss=False):
        key = ("taper", taper, int(n_window), int(n_norm), periodic_dpss,
               tuple(sorted((taper_opt or {}).items())))
        return self._cached(self._lru, MAX_TABLES, key, lambda: self.to_device(
            hm.normalized_tapers(taper, n_window, n_norm, taper_opt, periodic_dpss).astype(np.float32)))

    def scratch(self, name, shape, dtype):
        """Grow-only, reused device scratch tensor (e.g. the spectra handed from K1 to K2)."""
        n = int(np.prod(shape))
        key = ("scratch", name, dtype)
        with self._lock:
            buf = self._tables.get(key)
            if buf is None or buf.numel() < n:
                self._tables[key] = None
                buf = torch.empty(n, dtype=dtype, device=self.tdev)
                self._tables[key] = buf
        return buf[:n].view(shape)

    def index_table(self, idx, cache=True):
        """int32 index list on the device; `cache=False` for per-call lists (gather tables of single trials)."""
        if idx is None:
            return None
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        if not cache or idx.size > 65536:
            return torch.from_numpy(idx).to(self.tdev)
        return self._cached(self._lru, MAX_TABLES, ("idx", idx.tobytes()), lambda: torch.from_numpy(idx).to(self.tdev))

    # ------------------------------------------------------------------ K1: mtmfft
    def mtmfft(self, x, tapers, nfft, scale, polyremoval=-1, demean_taper=False, freq_idx=None,
               output="fourier", keeptapers=True, out=None, freq_major=False, chan_amax=None):
        """
        x [B, N, C] float32 CUDA (trial stride may exceed N*C), tapers [K, N] float32 CUDA.
        Returns out [B, Kout, nF, C] (default) or, with `freq_major`, [nF, B*Kout, C] -- the
        layout the cross-spectral contraction consumes.
        """
        tstride = _trial_layout(x)
        B, N, Cn = x.shape
        K = tapers.shape[0]
        assert tapers.shape[1] == N and tapers.is_contiguous()
        planar = output == "fourier_planar"
        kind = 8 if planar else hm.out_kind(output)
        fidx = self.index_table(freq_idx)
        nF = (nfft // 2 + 1) if fidx is None else fidx.numel()
        Kout = K if keeptapers else 1
        dt = _CDTYPE[kind == 2]
        if planar:
            # float32 [nF, R, 2, C]: the operand layout of the tcgen05 cross-spectral kernel
            assert freq_major and keeptapers
            shape = (nF, B * Kout, 2, Cn)
        else:
            shape = (nF, B * Kout, Cn) if freq_major else (B, Kout, nF, Cn)
        if out is None:
            out = torch.empty(shape, dtype=dt, device=self.tdev)
        assert out.is_cuda and out.dtype == dt and tuple(out.shape) == shape
        if planar:
            assert out[0].is_contiguous()
            so_freq, so_trial, so_taper = (out.stride(0) if nF > 1 else 0), Kout * 2 * Cn, 2 * Cn
        elif freq_major:
            # `out` may be a row-slice of a larger [nF, R, C] buffer (trial chunks)
            assert out[0].is_contiguous()
            so_freq, so_trial, so_taper = (out.stride(0) if nF > 1 else 0), Kout * Cn, Cn
        else:
            assert out.is_contiguous()
            so_trial, so_taper, so_freq = Kout * nF * Cn, nF * Cn, Cn
        esize = out.element_size()
        for b0 in range(0, B, MAX_LAUNCH_DIM):            # one launch covers at most 65535 trials (grid.z)
            nb = min(MAX_LAUNCH_DIM, B - b0)
            _lib.check(self.lib.spyb_mtmfft(
                x.data_ptr() + b0 * tstride * 4, nb, tstride, N, Cn, tapers.data_ptr(), K, int(nfft), float(scale),
                int(polyremoval), int(bool(demean_taper)), _ptr(fidx), nF, kind, int(bool(keeptapers)),
                out.data_ptr() + b0 * so_trial * esize, so_trial, so_taper, so_freq, _ptr(chan_amax), self.stream()))
        return out

    # ------------------------------------------------------------------ K4: mtmconvol
    def mtmconvol(self, x, tapers, nperseg, hop, frame_start0, n_frames, scale, polyremoval=-1,
                  freq_idx=None, output="fourier", keeptapers=True, out=None):
        """x [B, N, C] -> out [B, n_frames, Kout, nF, C]."""
        tstride = _trial_layout(x)
        B, N, Cn = x.shape
        K = tapers.shape[0]
        assert tapers.shape[1] == nperseg and tapers.is_contiguous()
        kind = hm.out_kind(output)
        fidx = self.index_table(freq_idx)
        nF = (nperseg // 2 + 1) if fidx is None else fidx.numel()
        Kout = K if keeptapers else 1
        shape = (B, n_frames, Kout, nF, Cn)
        dt = _CDTYPE[kind == 2]
        if out is None:
            out = torch.empty(shape, dtype=dt, device=self.tdev)
        else:
            assert out.is_cuda and out.dtype == dt and out.is_contiguous() and out.numel() == int(np.prod(shape))
        so_freq = Cn
        so_taper = nF * Cn
        so_frame = Kout * so_taper
        so_trial = n_frames * so_frame
        esize = out.element_size()
        # one launch covers at most 65535 frames (grid.y) x 65535 trials (grid.z): toi='all' on long trials has more
        for b0 in range(0, B, MAX_LAUNCH_DIM):
            nb = min(MAX_LAUNCH_DIM, B - b0)
            for f0 in range(0, max(n_frames, 1), MAX_LAUNCH_DIM):
                nf = min(MAX_LAUNCH_DIM, n_frames - f0)
                if nf <= 0:
                    break
                _lib.check(self.lib.spyb_mtmconvol(
                    x.data_ptr() + b0 * tstride * 4, nb, tstride, N, Cn, tapers.data_ptr(), K, int(nperseg), int(hop),
                    int(frame_start0 + f0 * hop), int(nf), float(scale), int(polyremoval), _ptr(fidx), nF, kind,
                    int(bool(keeptapers)), out.data_ptr() + (b0 * so_trial + f0 * so_frame) * esize,
                    so_trial, so_frame, so_taper, so_freq, self.stream()))
        return out

    # ------------------------------------------------------------------ K2 / K3
    def csd_accumulate(self, spectra, acc=None, alpha=1.0, beta=0.0, idx_i=None, idx_j=None, impl=0):
        """
        spectra [nF, R, C] complex64 (rows r = trial*K + taper), acc [nF, Ci, Cj] complex64:
        acc = beta*acc + alpha * sum_r X_r[si] conj(X_r[sj]).
        """
        assert spectra.is_cuda and spectra.dtype == torch.complex64 and spectra.dim() == 3
        nF, R, Cn = spectra.shape
        assert spectra[0].is_contiguous()
        ti, tj = self.index_table(idx_i), self.index_table(idx_j)
        Ci = Cn if ti is None else ti.numel()
        Cj = Cn if tj is None else tj.numel()
        if acc is None:
            acc = torch.empty((nF, Ci, Cj), dtype=torch.complex64, device=self.tdev)
            beta = 0.0
        assert acc.is_contiguous() and acc.shape == (nF, Ci, Cj) and acc.dtype == torch.complex64
        _lib.check(self.lib.spyb_csd_accumulate(
            spectra.data_ptr(), (spectra.stride(0) if nF > 1 else 0), Cn, R, nF, Cn,
            _ptr(ti), Ci, _ptr(tj), Cj, float(alpha), float(beta), acc.data_ptr(), int(impl), self.stream()))
        return acc

    def csd_planar_supported(self, n_chan):
        return bool(self.lib.spyb_csd_planar_supported(int(n_chan), 4, 4))

    def csd_accumulate_planar(self, planes, acc=None, alpha=1.0, beta=0.0):
        """
        planes [nF, R, 2, C] float32 (re / im planes per row, from `mtmfft(output="fourier_planar")`),
        acc [nF, C, C] complex64: acc = beta*acc + alpha * sum_r X_r X_r^H on the tcgen05 tensor cores.
        """
        assert planes.is_cuda and planes.dtype == torch.float32 and planes.dim() == 4 and planes.shape[2] == 2
        nF, R, _, Cn = planes.shape
        assert planes[0].is_contiguous()
        if acc is None:
            acc = torch.empty((nF, Cn, Cn), dtype=torch.complex64, device=self.tdev)
            beta = 0.0
        assert acc.is_contiguous() and acc.shape == (nF, Cn, Cn) and acc.dtype == torch.complex64
        _lib.check(self.lib.spyb_csd_accumulate_planar(
            planes.data_ptr(), (planes.stride(0) if nF > 1 else 2 * R * Cn), 2 * Cn, R, nF, Cn,
            float(alpha), float(beta), acc.data_ptr(), self.stream()))
        return acc

    def csd_tile_count(self, n_chan):
        return int(self.lib.spyb_csd_tile_count(int(n_chan)))

    def csd_accumulate_tiles(self, planes, owner_ptrs, f_begin, src_rank=0, alpha=1.0, beta=0.0, skip_own=False):
        """
        planes [nF, R, 2, C] float32 -> the upper 128x128 tiles of sum_r X_r X_r^H, frequency f written into the
        slot buffer of the rank owning f: `owner_ptrs[o]` is the (possibly peer-mapped) device address of rank o's
        buffer [n_src, f_begin[o+1]-f_begin[o], n_tiles, 128, 128] complex64, this rank fills source slot `src_rank`.
        `skip_own`: leave out the frequencies `src_rank` owns itself (they go through `csd_coherence_planar`
        with `add_slots` after the barrier).
        """
        import ctypes as C
        assert planes.is_cuda and planes.dtype == torch.float32 and planes.dim() == 4 and planes.shape[2] == 2
        nF, R, _, Cn = planes.shape
        assert planes[0].is_contiguous() and len(f_begin) == len(owner_ptrs) + 1
        n_own = len(owner_ptrs)
        ptrs = (C.c_void_p * n_own)(*[int(p) for p in owner_ptrs])
        fb = (C.c_int * (n_own + 1))(*[int(v) for v in f_begin])
        fn = self.lib.spyb_csd_accumulate_tiles_others if skip_own else self.lib.spyb_csd_accumulate_tiles
        _lib.check(fn(planes.data_ptr(), (planes.stride(0) if nF > 1 else 2 * R * Cn), 2 * Cn, R, nF, Cn,
                      float(alpha), float(beta), ptrs, fb, n_own, int(src_rank), self.stream()))

    def csd_coherence_planar(self, planes, output="abs", out=None, add_slots=None, skip_src=-1):
        """
        planes [nF, R, 2, C] float32 with ALL (trial, taper) rows -> coherency [nF, C, C] in one kernel: tcgen05
        contraction whose epilogue normalises, converts and mirrors (no cross-spectral matrix in memory).
        `add_slots` [n_src, nF, n_tiles, 128, 128] complex64: partial sums of other ranks for the same nF
        frequencies (tile-slot layout), added before the normalisation; source `skip_src` is not read.
        """
        assert planes.is_cuda and planes.dtype == torch.float32 and planes.dim() == 4 and planes.shape[2] == 2
        nF, R, _, Cn = planes.shape
        assert planes[0].is_contiguous()
        kind = hm.out_kind(output)
        if out is None:
            out = torch.empty((nF, Cn, Cn), dtype=_CDTYPE[kind == 2], device=self.tdev)
        assert out.is_contiguous() and out.numel() == nF * Cn * Cn and out.dtype == _CDTYPE[kind == 2]
        sx_f = planes.stride(0) if nF > 1 else 2 * R * Cn
        if add_slots is not None:
            assert add_slots.is_cuda and add_slots.dtype == torch.complex64 and add_slots.is_contiguous()
            assert tuple(add_slots.shape[1:]) == (nF, self.csd_tile_count(Cn), 128, 128)
            _lib.check(self.lib.spyb_csd_coherence_planar_slots(
                planes.data_ptr(), sx_f, 2 * Cn, R, nF, Cn, add_slots.data_ptr(), int(add_slots.shape[0]),
                int(skip_src), kind, out.data_ptr(), self.stream()))
            return out
        _lib.check(self.lib.spyb_csd_coherence_planar(
            planes.data_ptr(), sx_f, 2 * Cn, R, nF, Cn, kind, out.data_ptr(), self.stream()))
        return out

    def csd_normalize_tiles(self, slots, n_chan, output="abs", pre_scale=1.0, out=None):
        """slots [n_src, nF_loc, n_tiles, 128, 128] complex64 (local) -> coherency [nF_loc, C, C] of the summed slots."""
        assert slots.is_cuda and slots.dtype == torch.complex64 and slots.is_contiguous() and slots.dim() == 5
        n_src, nF = slots.shape[0], slots.shape[1]
        kind = hm.out_kind(output)
        if out is None:
            out = torch.empty((nF, n_chan, n_chan), dtype=_CDTYPE[kind == 2], device=self.tdev)
        assert out.is_contiguous() and out.numel() == nF * n_chan * n_chan and out.dtype == _CDTYPE[kind == 2]
        _lib.check(self.lib.spyb_csd_normalize_tiles(slots.data_ptr(), n_src, nF, int(n_chan), float(pre_scale), kind,
                                                     out.data_ptr(), self.stream()))
        return out

    def csd_normalize(self, csd, output="abs", pre_scale=1.0, out=None):
        """csd [..., C, C] complex64 -> coherency converted to `output` (same shape)."""
        assert csd.is_cuda and csd.dtype == torch.complex64 and csd.is_contiguous()
        Cn = csd.shape[-1]
        assert csd.shape[-2] == Cn
        n_mat = csd.numel() // (Cn * Cn)
        kind = hm.out_kind(output)
        if out is None:
            out = torch.empty(csd.shape, dtype=_CDTYPE[kind == 2], device=self.tdev)
        _lib.check(self.lib.spyb_csd_normalize(csd.data_ptr(), n_mat, Cn, float(pre_scale), kind,
                                               out.data_ptr(), self.stream()))
        return out

    # ------------------------------------------------------------------ K5 / K6: wavelets, superlets
    def detrend(self, x, polyremoval):
        """Whole-trial detrend of x [B, N, C] (new tensor); polyremoval -1 / 0 / 1."""
        tstride = _trial_layout(x)
        B, N, Cn = x.shape
        out = torch.empty((B, N, Cn), dtype=torch.float32, device=self.tdev)
        _lib.check(self.lib.spyb_detrend(x.data_ptr(), B, tstride, N, Cn, int(polyremoval), out.data_ptr(),
                                         N * Cn, self.stream()))
        return out

    def conv_plan(self, key, n_samples, taps_per_scale, exponents):
        """
        Device tables of a 'same'-convolution transform: `taps_per_scale[s][j]` are the sampled kernels of scale
        s (float64 / complex128 host arrays), `exponents[s][j]` their powers.  Cached under `key`.
        """
        key = ("conv", key, int(n_samples))

        def make():
            flat = [t for tl in taps_per_scale for t in tl]
            L = hm.conv_same_length(n_samples, flat)
            if L > MAX_LONG_FFT:
                raise _lib.SpybError(f"transform of {n_samples} samples needs a circular length of {L} > {MAX_LONG_FFT}")
            nS = len(taps_per_scale)
            if os.environ.get("SPYB_CWT_NO_SEGMENTS"):
                groups = [dict(s0=0, s1=nS, L=L, seg=False, V=n_samples, A=0, n_seg=1)]
            elif os.environ.get("SPYB_CWT_SEG_LENGTHS"):         # experiments: candidate block lengths
                cand = tuple(int(v) for v in os.environ["SPYB_CWT_SEG_LENGTHS"].split(","))
                groups = hm.conv_groups(n_samples, taps_per_scale, L, candidates=cand)
            else:
                groups = hm.conv_groups(n_samples, taps_per_scale, L)
            for g in groups:
                tls, els = taps_per_scale[g["s0"]:g["s1"]], exponents[g["s0"]:g["s1"]]
                max_fac = max(len(tl) for tl in tls)
                kern = np.zeros((len(tls), max_fac, g["L"]), dtype=np.complex64)
                expo = np.ones((len(tls), max_fac), dtype=np.float32)
                nfac = np.zeros(len(tls), dtype=np.int32)
                for si, (tl, el) in enumerate(zip(tls, els)):
                    nfac[si] = len(tl)
                    for j, (taps, e) in enumerate(zip(tl, el)):
                        kern[si, j] = (hm.conv_segment_spectrum(taps, g["L"], g["A"]) if g["seg"]
                                       else hm.conv_same_spectrum(taps, n_samples, g["L"]))
                        expo[si, j] = e
                g.update(max_fac=max_fac, kern=torch.from_numpy(kern).to(self.tdev),
                         expo=torch.from_numpy(expo).to(self.tdev), nfac=torch.from_numpy(nfac).to(self.tdev),
                         ones=torch.ones((1, g["L"]), dtype=torch.float32, device=self.tdev) if g["seg"] else None)
            return dict(L=L, n_scales=nS, groups=groups,
                        ones=torch.ones((1, n_samples), dtype=torch.float32, device=self.tdev))
        return self._cached(self._plans, MAX_PLANS, key, make)

    def cwt(self, x, plan, output="fourier", out=None):
        """
        x [B, N, C] float32 (already detrended / time-selected) -> [B, N, nScales, C]: per scale the product of
        powers of 'same' convolutions described by `plan` (see `conv_plan`).  Runs of scales with short kernels go
        through overlap-save blocks of the trial (`hostmath.conv_groups`), the others through one transform of the
        full padded length.
        """
        B, N, Cn = x.shape
        nS = plan["n_scales"]
        kind = hm.out_kind(output)
        dt = _CDTYPE[kind == 2]
        esize = 8 if kind == 2 else 4
        if out is None:
            out = torch.empty((B, N, nS, Cn), dtype=dt, device=self.tdev)
        assert out.is_contiguous() and out.dtype == dt and tuple(out.shape) == (B, N, nS, Cn)
        smem_max = self.lib.spyb_max_fft_len(1)
        full_spec = {}                                   # forward spectra of the whole trials, shared by the full-length runs

        def full_spectra(L):
            if L not in full_spec:
                full_spec[L] = self.mtmfft(x, plan["ones"], L, 1.0, polyremoval=-1, output="fourier", keeptapers=True)
            return full_spec[L]

        for g in plan["groups"]:
            L, nSg, s0 = g["L"], g["s1"] - g["s0"], g["s0"]
            if not g["seg"] and L > smem_max:
                # global-memory path (csrc/fft_long.cu): columns = (scale, channel), reference layouts on both sides
                xspec = full_spectra(L)
                direct = nSg == nS
                tmp = out if direct else torch.empty((B, N, nSg, Cn), dtype=dt, device=self.tdev)
                _lib.check(self.lib.spyb_cwt(xspec.data_ptr(), B, Cn, L, g["kern"].data_ptr(), g["expo"].data_ptr(),
                                             g["nfac"].data_ptr(), nSg, g["max_fac"], N, kind, 0, tmp.data_ptr(),
                                             self.stream()))
                if not direct:
                    out[:, :, s0:g["s1"]] = tmp
                continue
            # the transform kernel owns one channel of one scale per block: feed it channel-major spectra and let it
            # write time-contiguous rows, then transpose into the slice [time][scales of the run][channel] of the result
            nF = L // 2 + 1
            if g["seg"]:
                n_seg, V = g["n_seg"], g["V"]
                xspec = self.mtmconvol(x, g["ones"], L, V, -g["A"], n_seg, 1.0, polyremoval=-1, output="fourier",
                                       keeptapers=True)                      # [B, n_seg, 1, nF, C]
            else:
                n_seg, V = 1, N
                xspec = full_spectra(L)
            nb = B * n_seg
            step = max(1, min(B, MAX_LAUNCH_DIM // n_seg))                   # (trial, segment) pairs per launch
            xs_t = self.scratch("cwt_xspec_t", (nb, Cn, nF), torch.complex64)
            out_t = self.scratch("cwt_out_t" + ("c" if kind == 2 else "f"), (min(step, B) * n_seg, nSg * Cn, V), dt)
            for b0 in range(0, B, step):
                nbb = min(step, B - b0) * n_seg
                xin = xspec.view(nb, nF, Cn)[b0 * n_seg:]
                _lib.check(self.lib.spyb_transpose(xin.data_ptr(), xs_t.data_ptr(), nbb, nF, Cn, 8, self.stream()))
                _lib.check(self.lib.spyb_cwt(xs_t.data_ptr(), nbb, Cn, L, g["kern"].data_ptr(), g["expo"].data_ptr(),
                                             g["nfac"].data_ptr(), nSg, g["max_fac"], V, kind, 1, out_t.data_ptr(),
                                             self.stream()))
                dst = out.data_ptr() + (b0 * N * nS * Cn + s0 * Cn) * esize
                _lib.check(self.lib.spyb_transpose_place(out_t.data_ptr(), dst, nbb // n_seg, n_seg, nSg * Cn, V, esize,
                                                         N * nS * Cn, V * nS * Cn, nS * Cn, N, self.stream()))
        return out

    def gather_rows(self, src, idx):
        """src [B, R, ...] float32 / complex64 -> [B, len(idx), ...] (rows idx of every trial)."""
        assert src.is_contiguous()
        B, R = src.shape[:2]
        tidx = self.index_table(idx, cache=False)
        out = torch.empty((B, tidx.numel()) + tuple(src.shape[2:]), dtype=src.dtype, device=self.tdev)
        words = (2 if src.dtype == torch.complex64 else 1)
        row = int(np.prod(src.shape[2:])) * words
        _lib.check(self.lib.spyb_gather_rows(src.data_ptr(), B, R * row, tidx.data_ptr(), tidx.numel(), row,
                                             out.data_ptr(), self.stream()))
        return out

    # ------------------------------------------------------------------ K7-K9: Granger causality (float64)
    def _workspace(self, nbytes):
        """Grow-only device scratch shared by the Granger kernels."""
        with self._lock:
            ws = self._tables.get("workspace")
            if ws is None or ws.numel() < nbytes:
                self._tables["workspace"] = None
                ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.tdev)
                self._tables["workspace"] = ws
        return ws

    def regularize_csd(self, csd, cond_max=1e3, eps_max=1e-3, n_steps=15):
        """
        csd [nF, C, C] complex64 CUDA -> (csd + eps*I as complex128, factor, initial condition number);
        factor is 0 / eps / -1 exactly as `regularize_csd` reports it (wilson_sf.py:243-254).
        """
        import ctypes as C
        assert csd.is_cuda and csd.dtype == torch.complex64 and csd.dim() == 3 and csd.is_contiguous()
        nF, Cn, _ = csd.shape
        out = torch.empty((nF, Cn, Cn), dtype=torch.complex128, device=self.tdev)
        nbytes = self.lib.spyb_regularize_workspace_bytes(nF, Cn)
        ws = self._workspace(nbytes)
        eps, cond0 = C.c_double(0.0), C.c_double(0.0)
        _lib.check(self.lib.spyb_regularize_csd(csd.data_ptr(), nF, Cn, float(cond_max), float(eps_max),
                                                int(n_steps), out.data_ptr(), C.byref(eps), C.byref(cond0),
                                                ws.data_ptr(), ws.numel(), self.stream()))
        factor = eps.value
        if factor == 0.0 or factor == -1.0:
            factor = int(factor)
        return out, factor, cond0.value

    def wilson_sf(self, csd, n_iter=100, rtol=1e-6, slab=None, exchange=None):
        """
        csd [nF, C, C] complex128 CUDA (one-sided) -> (H [nF, C, C] complex128, Sigma [C, C] float64,
        converged, err, iterations)  -- wilson_sf.py:16-120.

        With `slab = (f_lo, f_hi)` and `exchange(what, buf_ptr, row_bytes, n_rows) -> int` (see
        `spyb_wilson_sharded` in include/spyb200.h; `distributed.WilsonExchange` is the torch.distributed one) only
        the slab is factorised and H is valid for the slab's rows only.
        """
        import ctypes as C
        assert csd.is_cuda and csd.dtype == torch.complex128 and csd.dim() == 3 and csd.is_contiguous()
        nF, Cn, _ = csd.shape
        H = torch.empty((nF, Cn, Cn), dtype=torch.complex128, device=self.tdev)
        Sigma = torch.empty((Cn, Cn), dtype=torch.float64, device=self.tdev)
        nbytes = self.lib.spyb_wilson_workspace_bytes(nF, Cn)
        ws = self._workspace(nbytes)
        conv, iters, err = C.c_int(0), C.c_int(0), C.c_double(0.0)
        if exchange is None:
            _lib.check(self.lib.spyb_wilson(csd.data_ptr(), nF, Cn, int(n_iter), float(rtol), H.data_ptr(),
                                            Sigma.data_ptr(), C.byref(conv), C.byref(err), C.byref(iters),
                                            ws.data_ptr(), ws.numel(), self.stream()))
        else:
            failure = []

            def _cb(_ctx, what, buf, row_bytes, n_rows):
                try:
                    return int(exchange(int(what), int(buf), int(row_bytes), int(n_rows)) or 0)
                except Exception as exc:      # noqa: BLE001  (must not propagate through the C frame)
                    failure.append(exc)
                    return 1
            cb = _lib.EXCHANGE_FN(_cb)
            rc = self.lib.spyb_wilson_sharded(csd.data_ptr(), nF, Cn, int(n_iter), float(rtol), H.data_ptr(),
                                              Sigma.data_ptr(), C.byref(conv), C.byref(err), C.byref(iters),
                                              ws.data_ptr(), ws.numel(), int(slab[0]), int(slab[1]), cb, None,
                                              self.stream())
            if failure:
                raise failure[0]
            _lib.check(rc)
        return H, Sigma, bool(conv.value), err.value, iters.value

    def granger(self, csd, H, Sigma):
        """granger.py:10-79 on CUDA tensors -> [nF, C, C] float32."""
        assert csd.dtype == torch.complex128 and H.dtype == torch.complex128 and Sigma.dtype == torch.float64
        assert csd.is_contiguous() and H.is_contiguous() and Sigma.is_contiguous()
        nF, Cn, _ = csd.shape
        out = torch.empty((nF, Cn, Cn), dtype=torch.float32, device=self.tdev)
        _lib.check(self.lib.spyb_granger(csd.data_ptr(), H.data_ptr(), Sigma.data_ptr(), nF, Cn, out.data_ptr(),
                                         self.stream()))
        return out

    def sum_trials(self, x, acc=None, alpha=1.0, beta=0.0):
        """
        x [B, ...] float32 / complex64 (dense trailing dims) -> acc [...] = beta*acc + alpha * sum_b x[b]: the
        runtime's `target[()] += res` over a batch (computational_routine.py:1025) and, with alpha = 1/nTrials,
        its final division (:1030-1032).
        """
        assert x.is_cuda and x.dim() >= 2 and x[0].is_contiguous()
        B = x.shape[0]
        words = 2 if x.dtype == torch.complex64 else 1
        n = int(np.prod(x.shape[1:])) * words
        stride = (x.stride(0) if B > 1 else int(np.prod(x.shape[1:]))) * words
        if acc is None:
            acc = torch.empty(x.shape[1:], dtype=x.dtype, device=self.tdev)
            beta = 0.0
        assert acc.is_contiguous() and acc.dtype == x.dtype and tuple(acc.shape) == tuple(x.shape[1:])
        if n % 4 or stride % 4:      # odd sizes: rare (needs a channel count that is not a multiple of 4)
            red = x.sum(dim=0) * alpha
            acc.copy_(red if beta == 0.0 else acc * beta + red)
            return acc
        _lib.check(self.lib.spyb_sum_trials(x.data_ptr(), B, stride, n, float(alpha), float(beta), acc.data_ptr(),
                                            self.stream()))
        return acc

    # ------------------------------------------------------------------ jackknife / PPC / cross-covariance pieces
    @staticmethod
    def _words(t):
        return t.numel() * (2 if t.dtype == torch.complex64 else 1)

    def axpby(self, x, y, a, b, out=None):
        """out = a*x + b*y element-wise (float32 / complex64 with real factors); y may be None."""
        assert x.is_cuda and x.is_contiguous() and x.dtype in (torch.float32, torch.complex64)
        if out is None:
            out = torch.empty_like(x)
        assert out.is_contiguous() and out.dtype == x.dtype and out.numel() == x.numel()
        if y is not None:
            assert y.is_contiguous() and y.dtype == x.dtype and y.numel() == x.numel()
        _lib.check(self.lib.spyb_axpby(x.data_ptr(), _ptr(y), float(a), float(b), out.data_ptr(), self._words(x),
                                       self.stream()))
        return out

    def sqdev_accumulate(self, avg, x, var):
        """var += |avg - x|^2 (var float32, avg / x float32 or complex64)."""
        assert avg.is_contiguous() and x.is_contiguous() and var.is_contiguous() and var.dtype == torch.float32
        assert avg.dtype == x.dtype and avg.numel() == x.numel() == var.numel()
        _lib.check(self.lib.spyb_sqdev_accumulate(avg.data_ptr(), x.data_ptr(), var.data_ptr(), x.numel(),
                                                  int(x.dtype == torch.complex64), self.stream()))
        return var

    def unit_accumulate(self, z, acc, first):
        """acc (+)= z / |z| for a complex64 array (single-trial cross spectra)."""
        assert z.is_contiguous() and acc.is_contiguous() and z.dtype == acc.dtype == torch.complex64
        _lib.check(self.lib.spyb_unit_accumulate(z.data_ptr(), acc.data_ptr(), z.numel(), int(bool(first)), self.stream()))
        return acc

    def ppc_finish(self, acc, n_trials):
        out = torch.empty(acc.shape, dtype=torch.float32, device=self.tdev)
        _lib.check(self.lib.spyb_ppc_finish(acc.data_ptr(), out.data_ptr(), acc.numel(), int(n_trials), self.stream()))
        return out

    def cross_covariance(self, x, polyremoval=0, norm=False):
        """
        x [N, C] float32 CUDA (one trial) -> [nLags, C, C] float32: single-trial cross-covariance of
        `cross_covariance_cF` (ST_compRoutines.py:465-584) by FFT: one spectrum per channel, kernel spectra of the
        time-reversed channels, C^2 inverse transforms in one launch of the wavelet kernel, lag selection.
        """
        N, Cn = x.shape
        n_lags = N // 2 if N % 2 == 0 else N // 2 + 1
        L = 16
        while L < 2 * N - 1:
            L *= 2
        if L > MAX_LONG_FFT:
            raise _lib.SpybError(f"cross-covariance of {N} samples needs a circular length of {L} > {MAX_LONG_FFT}")
        ones = self._cached(self._lru, MAX_TABLES, ("ones", N), lambda: torch.ones((1, N), dtype=torch.float32, device=self.tdev))
        xspec = self.mtmfft(x[None], ones, L, 1.0, polyremoval=polyremoval, output="fourier", keeptapers=True)  # [1,1,nF,C]
        nF = L // 2 + 1
        xs_t = self.scratch("xcov_xspec_t", (1, Cn, nF), torch.complex64)
        _lib.check(self.lib.spyb_transpose(xspec.data_ptr(), xs_t.data_ptr(), 1, nF, Cn, 8, self.stream()))
        kern = self.scratch("xcov_kern", (Cn, 1, L), torch.complex64)
        n_time = 2 * n_lags + 1
        shift = N - 1 - n_lags
        _lib.check(self.lib.spyb_xcov_kernel_spectra(xs_t.data_ptr(), Cn, L, N, shift, kern.data_ptr(), self.stream()))
        expo = self._cached(self._lru, MAX_TABLES, ("xcov_expo", Cn), lambda: torch.ones((Cn, 1), dtype=torch.float32, device=self.tdev))
        nfac = self._cached(self._lru, MAX_TABLES, ("xcov_nfac", Cn), lambda: torch.ones(Cn, dtype=torch.int32, device=self.tdev))
        corr = self.scratch("xcov_corr", (Cn * Cn, n_time), torch.float32)
        _lib.check(self.lib.spyb_cwt(xs_t.data_ptr(), 1, Cn, L, kern.data_ptr(), expo.data_ptr(), nfac.data_ptr(), Cn, 1,
                                     n_time, hm.out_kind("real"), 1, corr.data_ptr(), self.stream()))
        out = torch.empty((n_lags, Cn, Cn), dtype=torch.float32, device=self.tdev)
        _lib.check(self.lib.spyb_xcov_finish(corr.data_ptr(), xs_t.data_ptr(), Cn, N, n_lags, L, int(bool(norm)),
                                             out.data_ptr(), self.stream()))
        return out

    # ------------------------------------------------------------------ preprocessing
    def fir_same(self, x, taps, key, output="real"):
        """
        x [B, N, C] float32 -> [B, N, C]: scipy.signal.convolve(x, taps[:, None], 'same') per channel as an FFT
        convolution on the wavelet kernel (one 'scale'); complex taps + `output` give analytic-signal conversions.
        """
        B, N, Cn = x.shape
        plan = self.conv_plan(("fir", key), N, [[np.asarray(taps)]], [[1.0]])
        return self.cwt(x, plan, output=output)[:, :, 0, :]

    def sosfilt(self, x, sos, twopass):
        """x [B, N, C] float32 -> float32: scipy.signal.sosfilt (twopass=False) / sosfiltfilt (True), float64 recursion."""
        import ctypes as C
        tstride = _trial_layout(x)
        B, N, Cn = x.shape
        sos = np.ascontiguousarray(sos, dtype=np.float64)
        S = sos.shape[0]
        out = torch.empty((B, N, Cn), dtype=torch.float32, device=self.tdev)
        edge, zi, scratch = 0, None, None
        if twopass:
            edge, zi = hm.sosfiltfilt_plan(sos)
            scratch = self.scratch("sos_fwd", (B, N + 2 * edge, Cn), torch.float64)
        zi_p = zi.ctypes.data_as(C.POINTER(C.c_double)) if zi is not None else None
        _lib.check(self.lib.spyb_sosfilt(x.data_ptr(), B, tstride, N, Cn, sos.ctypes.data_as(C.POINTER(C.c_double)), S,
                                         zi_p, int(edge), int(bool(twopass)), _ptr(scratch), out.data_ptr(), self.stream()))
        return out

    def resample_poly(self, x, h, up, down, first_row, n_out):
        """x [B, N, C] float32 -> [B, n_out, C]: rows first_row .. first_row + n_out of upfirdn(h, x, up, down)."""
        tstride = _trial_layout(x)
        B, N, Cn = x.shape
        hd = torch.from_numpy(np.ascontiguousarray(h, dtype=np.float64)).to(self.tdev)
        out = torch.empty((B, n_out, Cn), dtype=torch.float32, device=self.tdev)
        _lib.check(self.lib.spyb_upfirdn(x.data_ptr(), B, tstride, N, Cn, hd.data_ptr(), hd.numel(), int(up), int(down),
                                         int(first_row), int(n_out), out.data_ptr(), self.stream()))
        return out

    def standardize(self, x):
        tstride = _trial_layout(x)
        B, N, Cn = x.shape
        out = torch.empty((B, N, Cn), dtype=torch.float32, device=self.tdev)
        _lib.check(self.lib.spyb_standardize(x.data_ptr(), B, tstride, N, Cn, out.data_ptr(), self.stream()))
        return out

    def rectify(self, x):
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        out = torch.empty_like(x)
        _lib.check(self.lib.spyb_rectify(x.data_ptr(), out.data_ptr(), x.numel(), self.stream()))
        return out

    def csd_mirror_upper(self, csd):
        """In place on complex64 [nF, C, C]: lower triangle <- conj(upper), real diagonal (exact Hermitian symmetry
        after a reduction over ranks)."""
        assert csd.is_cuda and csd.dtype == torch.complex64 and csd.dim() == 3 and csd.is_contiguous()
        assert csd.shape[1] == csd.shape[2]
        for f0 in range(0, csd.shape[0], MAX_LAUNCH_DIM):
            nf = min(MAX_LAUNCH_DIM, csd.shape[0] - f0)
            _lib.check(self.lib.spyb_csd_mirror_upper(csd[f0:].data_ptr(), nf, csd.shape[1], self.stream()))
        return csd

    def scale_(self, t, s):
        """In-place t *= s for float32 / complex64 CUDA tensors."""
        assert t.is_cuda and t.is_contiguous()
        n = t.numel() * (2 if t.dtype == torch.complex64 else 1)
        _lib.check(self.lib.spyb_scale(t.data_ptr(), n, float(s), self.stream()))
        return t


_engines = {}


def get_engine(device=None):
    """Process-wide engine per device; the device defaults to LOCAL_RANK / SPYB_DEVICE / 0."""
    import os
    if device is None:
        device = int(os.environ.get("SPYB_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]
