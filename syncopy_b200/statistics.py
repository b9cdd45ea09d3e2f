"""
Trial statistics over single-trial cross spectra on the GPU -- SURVEY 8(f) rows 2 and 3:

    jackknife replicates, bias and variance     syncopy/statistics/jackknifing.py:14-184, used by
                                                syncopy/connectivity/connectivity_analysis.py:601-606,736-757
    pairwise phase consistency                  ST_compRoutines.py:158-233 + connectivity_analysis.py:624-667

The reference materialises every single-trial cross spectrum (1 GB per trial at 256 channels x 2049 bins) in an
HDF5 file, writes the T leave-one-out averages into a second file and runs the averaged routine on each.  Here the
single-trial cross spectra are regenerated from the trial's spectra when they are needed (one small contraction per
trial), so only a handful of [nFreq, C, C] arrays are resident whatever the number of trials; the arithmetic of each
replicate -- `(T * avg - x_k) / (T - 1)`, the averaged routine, `bias = (T-1) (mean - direct)`,
`var = (T-1) sum |mean - replicate|^2` -- is the reference's.  Multi-GPU: replicas only (every replicate needs all
trials' sum; shard the replicate index if needed).
"""
import numpy as np
import torch

from . import batched
from . import hostmath as hm
from .engine import get_engine


class _TrialCsd:
    """Single-trial cross spectra CS_k [nFreq, C, C] (complex64) on demand from resident per-trial spectra."""

    def __init__(self, eng, trials, samplerate, nSamples, foi, taper, taper_opt, demean_taper, polyremoval):
        self.eng = eng
        x = batched._device_trials(eng, trials)
        self.T, n_sig, self.C = x.shape
        nfft = n_sig if nSamples is None else int(nSamples)
        self.freqs, fidx = batched._freq_selection(nfft, samplerate, foi)
        tapers = eng.taper_table(taper, n_sig, nfft, taper_opt)
        self.K = tapers.shape[0]
        # frequency-major spectra of all trials: [nF, T*K, C] complex64 (K1, one launch)
        self.spectra = eng.mtmfft(x, tapers, nfft, hm.mtmfft_scale(n_sig, nfft), polyremoval=hm.polyremoval_code(polyremoval),
                                  demean_taper=demean_taper, freq_idx=fidx, output="fourier", keeptapers=True,
                                  freq_major=True)
        self.nF = self.spectra.shape[0]

    def trial(self, k, out=None):
        return self.eng.csd_accumulate(self.spectra[:, k * self.K:(k + 1) * self.K], acc=out, alpha=1.0 / self.K,
                                       beta=0.0, impl=1)

    def mean(self):
        """summary_stats.py:408-428: sum over trials / T."""
        return self.eng.csd_accumulate(self.spectra, alpha=1.0 / (self.K * self.T), impl=1)


class _ResidentCsd:
    """Single-trial cross spectra handed in as a stack [T, nF, C, C] complex64 (what the reference's `jack_in` is)."""

    def __init__(self, eng, stack):
        self.eng = eng
        x = stack if isinstance(stack, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(stack, dtype=np.complex64))
        self.stack = x.to(eng.tdev).contiguous()
        self.T, self.nF, self.C = self.stack.shape[0], self.stack.shape[1], self.stack.shape[2]
        self.freqs = None

    def trial(self, k, out=None):
        if out is None:
            return self.stack[k]
        out.copy_(self.stack[k])
        return out

    def mean(self):
        return self.eng.sum_trials(self.stack, alpha=1.0 / self.T)


def _jackknife(eng, src, averaged_routine, result_dtype):
    """Generic loop of connectivity_analysis.py:736-757 around an `averaged_routine(csd [nF,C,C]) -> [nF,C,C]`."""
    T = src.T
    if T < 2:
        raise ValueError("jackknife replicates with at least 2 trials")
    avg = src.mean()
    direct = averaged_routine(avg)
    cs = torch.empty_like(avg)
    loo = torch.empty_like(avg)
    jack_sum = None
    # pass 1: mean of the replicates (jackknifing.py:147)
    for k in range(T):
        src.trial(k, out=cs)
        eng.axpby(avg, cs, T / (T - 1.0), -1.0 / (T - 1.0), out=loo)          # (T*avg - x_k) / (T-1)   (:80-85)
        rep = averaged_routine(loo)
        jack_sum = eng.sum_trials(rep[None], acc=jack_sum, alpha=1.0, beta=0.0 if jack_sum is None else 1.0)
    jack_avg = eng.scale_(jack_sum, 1.0 / T)
    bias = eng.axpby(jack_avg, direct, float(T - 1), -float(T - 1))            # (T-1) (jack_avg - direct)   (:160)
    # pass 2: variance around the replicate mean, as the reference accumulates it (:164-170)
    var = torch.zeros(direct.shape, dtype=torch.float32, device=eng.tdev)
    for k in range(T):
        src.trial(k, out=cs)
        eng.axpby(avg, cs, T / (T - 1.0), -1.0 / (T - 1.0), out=loo)
        eng.sqdev_accumulate(jack_avg, averaged_routine(loo), var)
    eng.scale_(var, float(T - 1))
    return direct[None], bias[None], var[None]


def jackknife_coherence(trials, samplerate=1, nSamples=None, foi=None, taper="hann", taper_opt=None, polyremoval=0,
                        output="abs", to_host=False, engine=None):
    """
    `connectivityanalysis(method='coh', jackknife=True)`: (coherence [1,nF,C,C], jack_bias, jack_var, freqs).
    """
    eng = engine or get_engine()
    src = _TrialCsd(eng, trials, samplerate, nSamples, foi, taper, taper_opt, False, polyremoval)
    res = _jackknife(eng, src, lambda c: eng.csd_normalize(c[None], output=output)[0], None)
    if to_host:
        res = tuple(r.cpu().numpy() for r in res)
    return res + (src.freqs,)


def _granger_routine(eng, rtol, nIter, cond_max):
    def routine(csd):
        reg, _, _ = eng.regularize_csd(csd, cond_max=cond_max, eps_max=1e-1)
        H, Sigma, _, _, _ = eng.wilson_sf(reg, n_iter=nIter, rtol=rtol)
        return eng.granger(reg, H, Sigma)
    return routine


def jackknife_csd(single_trial_csd, method="coh", output="abs", rtol=5e-6, nIter=100, cond_max=1e4, to_host=False,
                  engine=None):
    """
    The jackknife of connectivity_analysis.py:601-606,736-757 on resident single-trial cross spectra
    [T, nF, C, C] complex64 (the reference's `jack_in`): (direct estimate, bias, variance) of coherence or Granger.
    """
    eng = engine or get_engine()
    src = _ResidentCsd(eng, single_trial_csd)
    if method == "coh":
        res = _jackknife(eng, src, lambda c: eng.csd_normalize(c[None], output=output)[0], None)
    elif method == "granger":
        res = _jackknife(eng, src, _granger_routine(eng, rtol, nIter, cond_max), None)
    else:
        raise ValueError("method must be 'coh' or 'granger'")
    return tuple(r.cpu().numpy() for r in res) if to_host else res


def jackknife_granger(trials, samplerate=1, nSamples=None, foi=None, taper="hann", taper_opt=None, polyremoval=0,
                      rtol=5e-6, nIter=100, cond_max=1e4, to_host=False, engine=None):
    """`connectivityanalysis(method='granger', jackknife=True)`: T + 1 Wilson factorisations, one after the other."""
    eng = engine or get_engine()
    src = _TrialCsd(eng, trials, samplerate, nSamples, foi, taper, taper_opt, True, polyremoval)
    res = _jackknife(eng, src, _granger_routine(eng, rtol, nIter, cond_max), None)
    if to_host:
        res = tuple(r.cpu().numpy() for r in res)
    return res + (src.freqs,)


def trial_avg_replicates(single_trials, engine=None):
    """jackknifing.py:14-108 on a resident stack [T, ...] (float32 / complex64 CUDA tensor or host array)."""
    eng = engine or get_engine()
    x = single_trials if isinstance(single_trials, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(single_trials))
    x = x.to(eng.tdev).contiguous()
    T = x.shape[0]
    avg = eng.sum_trials(x, alpha=1.0 / T)
    out = torch.empty_like(x)
    for k in range(T):
        eng.axpby(avg, x[k], T / (T - 1.0), -1.0 / (T - 1.0), out=out[k])
    return out


def bias_var(direct, replicates, engine=None):
    """jackknifing.py:111-184 on CUDA tensors: (bias, variance)."""
    eng = engine or get_engine()
    T = replicates.shape[0]
    if T <= 1:
        raise ValueError("jackknife replicates with at least 2 trials")
    jack_avg = eng.sum_trials(replicates, alpha=1.0 / T)
    bias = eng.axpby(jack_avg, direct.contiguous(), float(T - 1), -float(T - 1))
    var = torch.zeros(direct.shape, dtype=torch.float32, device=eng.tdev)
    for k in range(T):
        eng.sqdev_accumulate(jack_avg, replicates[k], var)
    return bias, eng.scale_(var, float(T - 1))


def ppc(trials, samplerate=1, nSamples=None, foi=None, taper="hann", taper_opt=None, polyremoval=0, to_host=False,
        engine=None):
    """
    `connectivityanalysis(method='ppc')`: pairwise phase consistency [1, nF, C, C] float32 over all trial pairs.
    The reference evaluates cos(angle(z_j conj z_k)) for the T(T-1)/2 pairs (one CR run per column, every pair read
    back from HDF5); with u_k = z_k / |z_k| the same average is (|sum_k u_k|^2 - T) / (T (T - 1)): one pass.
    """
    eng = engine or get_engine()
    src = _TrialCsd(eng, trials, samplerate, nSamples, foi, taper, taper_opt, False, polyremoval)
    if src.T < 2:
        raise ValueError("ppc needs at least two trials")
    cs = torch.empty((src.nF, src.C, src.C), dtype=torch.complex64, device=eng.tdev)
    acc = torch.empty_like(cs)
    for k in range(src.T):
        src.trial(k, out=cs)
        eng.unit_accumulate(cs, acc, first=(k == 0))
    out = eng.ppc_finish(acc, src.T)[None]
    return (out.cpu().numpy() if to_host else out), src.freqs
