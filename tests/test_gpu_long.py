"""
GPU parity of the global-memory transform path (syncopy_b200/csrc/fft_long.cu): FFT lengths beyond the shared-memory
kernels -- power-of-two lengths > 16384 and other lengths > 8192 -- which the reference takes like any other length
(scipy.fft.rfft in syncopy/specest/mtmfft.py:117-127, fftconvolve in syncopy/specest/wavelets/transform.py:88-108).
Mixed-radix lengths (prime factors <= 61) run as Stockham passes directly, anything else through Bluestein.
Tolerance: 1e-5 normwise against the oracle.
"""
import numpy as np
import pytest

from conftest import nerr
from oracle import connectivity as oc
from oracle import spectral as osp
from oracle import statistics as ost
from oracle import timefreq as otf
from oracle import synth

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _noise(n, c, seed):
    return np.random.default_rng(seed).normal(size=(n, c)).astype("f4")


@pytest.mark.parametrize("n,c,nfft", [
    (32768, 4, None),            # power of two: radix-16 passes + one radix-8
    (65536, 3, None),            # odd channel count (last pair half empty)
    (20000, 6, None),            # 2^5 5^4
    (18018, 2, None),            # 2 3^2 7 11 13: generic prime passes
    (10007, 3, None),            # prime > 8192: Bluestein on 32768
    (17389, 2, None),            # prime: Bluestein on 65536
    (9000, 5, 20000),            # zero padding up to a long transform
    (30011, 1, None),            # single channel, prime
])
def test_mtmfft_long_sizes_vs_oracle(engine, n, c, nfft):
    from syncopy_b200 import compute_functions as cf
    x = _noise(n, c, n + c)
    mk = dict(samplerate=1000., nSamples=nfft, taper="hann", taper_opt={})
    foi = np.fft.rfftfreq(nfft or n, 1e-3)
    got, _ = cf.mtmfft_cF(x.copy(), foi=foi, output="fourier", polyremoval=0, method_kwargs=mk)
    want, _ = osp.mtmfft_cF(x.copy(), foi=foi, output="fourier", polyremoval=0, method_kwargs=mk)
    assert got.shape == want.shape and got.dtype == want.dtype
    assert nerr(got, want) <= TOL


@pytest.mark.parametrize("output,keeptapers,pr", [("pow", False, 1), ("fourier", True, 0), ("abs", True, None),
                                                   ("fourier", False, 1)])
def test_mtmfft_long_tapers_outputs_foi(engine, output, keeptapers, pr):
    """DPSS tapers, taper mean, linear detrend, frequency selection (unsorted) and demean_taper at 40000 samples"""
    from syncopy_b200 import compute_functions as cf
    n = 40000
    x = _noise(n, 5, 11) + np.linspace(-0.01, 0.01, n, dtype="f4")[:, None]
    mk = dict(samplerate=2000., nSamples=None, taper="dpss", taper_opt={"NW": 3, "Kmax": 4}, demean_taper=True)
    foi = np.array([400.05, 10., 999.9, 0., 55.5, 1000.])
    got, _ = cf.mtmfft_cF(x.copy(), foi=foi, keeptapers=keeptapers, polyremoval=pr, output=output, method_kwargs=dict(mk))
    want, _ = osp.mtmfft_cF(x.copy(), foi=foi, keeptapers=keeptapers, polyremoval=pr, output=output, method_kwargs=dict(mk))
    assert got.shape == want.shape and got.dtype == want.dtype
    assert nerr(got, want) <= TOL


def test_mtmconvol_long_windows(engine):
    """sliding windows of 20000 samples (two 5-smooth transforms per trial, zero-padded edges)"""
    from syncopy_b200 import compute_functions as cf
    x = _noise(50000, 3, 5)
    mk = dict(samplerate=1000., nperseg=20000, noverlap=0, taper="hann", taper_opt={})
    foi = np.fft.rfftfreq(20000, 1e-3)[::40]
    kw = dict(equidistant=True, toi=0.0, foi=foi, keeptapers=False, polyremoval=0, output="pow")
    got = cf.mtmconvol_cF(x.copy(), slice(None), slice(None), method_kwargs=dict(mk), **kw)
    want = osp.mtmconvol_cF(x.copy(), slice(None), slice(None), method_kwargs=dict(mk), **kw)
    assert got.shape == want.shape and got.dtype == want.dtype
    assert nerr(got, want) <= TOL


def test_coherence_of_long_trials(engine):
    """planar spectra from the long path feeding the tcgen05 contraction: 6 trials x 20000 samples x 128 channels"""
    from syncopy_b200 import batched
    trials = synth.white_noise(6, 20000, 128)
    coh, freqs = batched.coherence(trials, 1000., taper="hann", polyremoval=0, to_host=True)
    # bin 0 is left out: after de-meaning it holds rounding noise only, and the coherence of noise with noise is
    # not a reproducible number (the long path sums the mean in float64, the reference in float32)
    fsel = np.array([1, 2, 77, 5000, 9999, 10000])
    acc = None
    for t in trials:
        specs, _ = osp.mtmfft(osp.detrend_trial(np.array(t), 0), 1000., None, "hann", None, False)
        s = specs[:, fsel, :]
        cs = (s[:, :, :, None] * s[:, :, None, :].conj()).mean(axis=0)
        acc = cs if acc is None else acc + cs
    want = oc.normalize_csd((acc / len(trials))[None], "abs")
    assert coh.shape == (1, 10001, 128, 128)
    assert nerr(coh[:, fsel], want) <= TOL
    assert np.isfinite(coh).all() and coh.max() <= 1 + 1e-6


@pytest.mark.parametrize("n,c,output", [(20000, 3, "pow"), (40000, 2, "fourier")])
def test_wavelet_long_trials(engine, n, c, output):
    from syncopy_b200 import compute_functions as cf
    from syncopy_b200 import hostmath as hm
    fs = 1000.
    x = _noise(n, c, n)
    wav_o, wav_g = otf.Morlet(6), hm.Morlet(6)
    foi = np.array([2., 11., 60., 240.])
    scales = wav_o.scale_from_period(1 / foi)
    kw = dict(toi=None, polyremoval=0, output=output)
    got = cf.wavelet_cF(x.copy(), slice(None), slice(None),
                        method_kwargs=dict(samplerate=fs, scales=scales, wavelet=wav_g), **kw)
    want = otf.wavelet_cF(x.copy(), slice(None), slice(None),
                          method_kwargs=dict(samplerate=fs, scales=scales, wavelet=wav_o), **kw)
    assert got.shape == want.shape and got.dtype == want.dtype
    assert nerr(got, want) <= TOL


@pytest.mark.parametrize("adaptive", [False, True])
def test_superlet_long_trials(engine, adaptive):
    from syncopy_b200 import compute_functions as cf
    fs, n = 1000., 18000
    x = _noise(n, 2, 9)
    foi = np.array([8., 30., 90.])
    scales = 1 / (2 * np.pi * foi)
    mk = dict(samplerate=fs, scales=scales, order_max=5, order_min=1, c_1=3, adaptive=adaptive)
    got = cf.superlet_cF(x.copy(), slice(None), slice(None), polyremoval=0, output="pow", method_kwargs=dict(mk))
    want = otf.superlet_cF(x.copy(), slice(None), slice(None), polyremoval=0, output="pow", method_kwargs=dict(mk))
    assert got.shape == want.shape and nerr(got, want) <= TOL


def test_cross_covariance_long(engine):
    from syncopy_b200 import compute_functions as cf
    x = _noise(9000, 3, 2)
    x[:, 1:] += 0.6 * np.roll(x[:, :1], 5, axis=0)
    got, lags = cf.cross_covariance_cF(x.copy(), samplerate=250., polyremoval=0, norm=True, fullOutput=True)
    want, lags0 = ost.cross_covariance_cF(x.copy(), samplerate=250., polyremoval=0, norm=True, fullOutput=True)
    assert got.shape == want.shape and np.array_equal(lags, lags0)
    assert nerr(got, want) <= TOL
