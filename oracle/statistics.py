"""
TEST INFRASTRUCTURE (see oracle/__init__.py) -- CPU restatement of the "next" rows of SURVEY 8(f):

    syncopy/statistics/jackknifing.py:14-108             -> trial_avg_replicates
    syncopy/statistics/jackknifing.py:111-184            -> bias_var
    syncopy/statistics/summary_stats.py:408-428          -> trial_mean (sequential `+=`, then `/=`)
    syncopy/connectivity/connectivity_analysis.py:601-606,736-757 -> jackknife_coherence / jackknife_granger
    syncopy/connectivity/ST_compRoutines.py:158-233      -> ppc_column_cF
    syncopy/connectivity/connectivity_analysis.py:624-667 -> ppc (pair loop, weights, normalisation)
    syncopy/connectivity/ST_compRoutines.py:465-584      -> cross_covariance_cF

The reference modules import h5py / the package itself and cannot be loaded here; `jackknifing.py` works on data
objects.  The arithmetic is restated on plain arrays [nTrials, ...], operation by operation and in the reference's
dtypes; `cross_covariance_cF` is additionally pinned against the reference's own formula by a live test that
executes the function body extracted from /root/reference (tests/test_oracle_vs_reference.py).
"""
import numpy as np
from scipy.signal import detrend, fftconvolve

from .connectivity import granger_cF, normalize_csd


def trial_mean(trials):
    """summary_stats.py:408-428: zeros of the data dtype, `+=` trial by trial, `/= nTrials`."""
    acc = np.zeros(trials[0].shape, dtype=trials[0].dtype)
    for t in trials:
        acc += t
    acc /= len(trials)
    return acc


def trial_avg_replicates(trials):
    """jackknifing.py:14-108: leave-one-out averages (nTrials * avg - trial_k) / (nTrials - 1), dtype of the data."""
    n = len(trials)
    avg = trial_mean(trials)
    reps = np.empty((n,) + avg.shape, dtype=avg.dtype)
    for k in range(n):
        loo = n * avg - trials[k]
        loo /= n - 1
        reps[k] = loo
    return reps


def bias_var(direct, replicates):
    """jackknifing.py:111-184: bias = (n-1) (mean(replicates) - direct); var = (n-1) sum |mean - replicate|^2 (float32)."""
    n = len(replicates)
    jack_avg = trial_mean(replicates)
    prefac = n - 1
    prefac = prefac + 0j if np.issubdtype(direct.dtype, np.complexfloating) else prefac
    bias = prefac * (jack_avg - direct)
    var = np.zeros(direct.shape, dtype=np.float32)
    for loo in replicates:
        var += (np.abs(jack_avg - loo)) ** 2
    var *= n - 1
    return bias, var


def jackknife_coherence(single_trial_csd, output="abs"):
    """connectivity_analysis.py:601-606,736-757 with NormalizeCrossSpectra: (direct, bias, variance, replicates)."""
    direct = normalize_csd(trial_mean(single_trial_csd)[None], output)[0]
    reps_avg = trial_avg_replicates(single_trial_csd)
    reps = np.stack([normalize_csd(r[None], output)[0] for r in reps_avg])
    bias, var = bias_var(direct, reps)
    return direct, bias, var, reps


def jackknife_granger(single_trial_csd, **kw):
    direct = granger_cF(trial_mean(single_trial_csd)[None], **kw)[0][0]
    reps_avg = trial_avg_replicates(single_trial_csd)
    reps = np.stack([granger_cF(r[None], **kw)[0][0] for r in reps_avg])
    bias, var = bias_var(direct, reps)
    return direct, bias, var, reps


def ppc_column_cF(cross_spectrum, cross_spectrum2):
    """ST_compRoutines.py:158-233 (the second trial is read from HDF5 there)."""
    return np.cos(np.angle(cross_spectrum * cross_spectrum2.conj()))


def ppc(single_trial_csd):
    """connectivity_analysis.py:624-667: float32 accumulator, upper-triangle weights, final 2 / nTrials."""
    n = len(single_trial_csd)
    acc = np.zeros(single_trial_csd[0].shape, dtype=np.float32)
    weights = np.arange(1, n) / (n - 1)
    for trl_idx in range(1, n):
        pairs = [ppc_column_cF(single_trial_csd[j], single_trial_csd[trl_idx]) for j in range(trl_idx)]
        acc += trial_mean(pairs) * weights[trl_idx - 1]
    acc *= 2 / n
    return acc


def cross_covariance_cF(trl_dat, samplerate=1, polyremoval=0, timeAxis=0, norm=False, fullOutput=False,
                        chunkShape=None, noCompute=False):
    """ST_compRoutines.py:465-584."""
    dat = trl_dat.T if timeAxis != 0 else trl_dat
    n, n_chan = dat.shape
    lags = np.arange(0, n // 2) if n % 2 == 0 else np.arange(0, n // 2 + 1)
    lags = lags * 1 / samplerate
    out_shape = (len(lags), 1, n_chan, n_chan)
    if noCompute:
        return out_shape, np.float32
    if polyremoval == 0:
        dat = detrend(dat, type="constant", axis=0, overwrite_data=True)
    elif polyremoval == 1:
        dat = detrend(dat, type="linear", axis=0, overwrite_data=True)
    norm_overlap = np.arange(n, n // 2, step=-1)
    CC = np.empty(out_shape)
    for i in range(n_chan):
        for j in range(i + 1):
            cc12 = fftconvolve(dat[:, i], dat[::-1, j], mode="same")
            CC[:, 0, i, j] = cc12[n // 2:] / norm_overlap
            if i != j:
                cc21 = cc12[::-1]
                CC[:, 0, j, i] = cc21[n // 2:] / norm_overlap
    if norm:
        stds = np.std(dat, axis=0)
        CC = CC / (stds[:, None] * stds[None, :])
    return (CC, lags) if fullOutput else CC
