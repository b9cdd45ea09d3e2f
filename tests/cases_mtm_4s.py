"""
GPU parity of the four-step tapered-FFT kernel (csrc/mtm_4s.cu: N = 4096, one taper, planar result, 32-channel
units, phase A / phase B in one persistent kernel with an L2-resident ring buffer) against the oracle: the headline
launch shape at small sizes, both detrending modes it takes (none / constant), the first-sample de-meaning trick
under large recording offsets, more units than one scheduling round and than the ring holds, trial strides, and
agreement with the interleaved output of the 8-channel kernel.
Reference: syncopy/specest/mtmfft.py:16-129, compRoutines.py:169-189.  Tolerance 1e-5 normwise.
"""
import numpy as np
import pytest
import torch

from conftest import nerr
from oracle import spectral as osp
from oracle import synth

pytestmark = pytest.mark.gpu
TOL = 1e-5
N = 4096


def _planar(engine, x, polyremoval, taper="hann", opt=None):
    from syncopy_b200 import hostmath as hm
    tapers = engine.taper_table(taper, N, N, opt)
    planes = engine.mtmfft(x, tapers, N, hm.mtmfft_scale(N, N), polyremoval=hm.polyremoval_code(polyremoval),
                           output="fourier_planar", keeptapers=True, freq_major=True)
    assert planes.shape == (N // 2 + 1, x.shape[0], 2, x.shape[2])
    return torch.complex(planes[:, :, 0, :], planes[:, :, 1, :]).cpu().numpy()           # [f, trial, c]


@pytest.mark.parametrize("polyremoval", [None, 0])
@pytest.mark.parametrize("n_trials,n_chan", [(5, 256), (16, 32), (3, 192)])
def test_vs_oracle(engine, polyremoval, n_trials, n_chan):
    trials = synth.white_noise(n_trials, N, n_chan) + np.float32(0.02)
    trials += np.linspace(-0.04, 0.05, n_chan, dtype="f4")[None, None, :]
    got = _planar(engine, torch.from_numpy(trials).to(engine.tdev), polyremoval)
    mk = dict(samplerate=1000., nSamples=None, taper="hann", taper_opt={})
    foi = np.fft.rfftfreq(N, 1e-3)
    for t in range(n_trials):
        want, _ = osp.mtmfft_cF(trials[t].copy(), foi=foi, keeptapers=True, polyremoval=polyremoval, output="fourier",
                                method_kwargs=mk)
        assert nerr(got[:, t, :], want[0, 0]) <= TOL


def test_large_offsets_vs_float64_detrend(engine):
    """recording offsets of 1000 standard deviations: the kernel subtracts the first sample before the taper and
    corrects with the taper's spectrum, so nothing cancels; the bar is the chain de-meaned in float64"""
    n_trials, n_chan = 16, 32
    trials = synth.white_noise(n_trials, N, n_chan)
    trials += (1000.0 * np.linspace(-1.0, 1.0, n_chan, dtype="f4"))[None, None, :]
    got = _planar(engine, torch.from_numpy(trials).to(engine.tdev), 0)
    mk = dict(samplerate=1000., nSamples=None, taper="hann", taper_opt={})
    foi = np.fft.rfftfreq(N, 1e-3)
    for t in (0, 7, 15):
        x64 = trials[t].astype(np.float64)
        exact_in = (x64 - x64.mean(axis=0, keepdims=True)).astype("f4")
        exact, _ = osp.mtmfft_cF(exact_in, foi=foi, keeptapers=True, polyremoval=None, output="fourier",
                                 method_kwargs=mk)
        assert nerr(got[:, t, :], exact[0, 0]) <= TOL


def test_many_units_ring_reuse_and_stride(engine):
    """more units than the ring buffer holds (3 rounds of 37 units on 148 SMs x 2 blocks): 60 trials x 8 groups = 480
    units, from a strided view of a larger buffer; spot-checked against the oracle and against the 8-channel kernel's
    interleaved output for every trial"""
    from syncopy_b200 import hostmath as hm
    rng = np.random.default_rng(11)
    big = torch.from_numpy(rng.normal(size=(60, N + 24, 256)).astype("f4")).to(engine.tdev)
    x = big[:, 8:8 + N]
    got = _planar(engine, x, 0, taper="dpss", opt={"NW": 2, "Kmax": 1})
    tapers = engine.taper_table("dpss", N, N, {"NW": 2, "Kmax": 1})
    inter = engine.mtmfft(x, tapers, N, hm.mtmfft_scale(N, N), polyremoval=0, output="fourier", keeptapers=True,
                          freq_major=True).cpu().numpy()                                   # [f, trial, c]
    assert inter.shape == got.shape
    assert nerr(got, inter) <= 5e-6
    xh = x.cpu().numpy()
    mk = dict(samplerate=1000., nSamples=None, taper="dpss", taper_opt={"NW": 2, "Kmax": 1})
    foi = np.fft.rfftfreq(N, 1e-3)
    for t in (0, 31, 59):
        want, _ = osp.mtmfft_cF(xh[t].copy(), foi=foi, keeptapers=True, polyremoval=0, output="fourier",
                                method_kwargs=mk)
        assert nerr(got[:, t, :], want[0, 0]) <= TOL


def test_repeatable_bitwise(engine):
    """fixed summation orders everywhere: two launches give identical bits"""
    trials = torch.from_numpy(synth.white_noise(20, N, 64)).to(engine.tdev)
    a = _planar(engine, trials, 0)
    b = _planar(engine, trials, 0)
    assert np.array_equal(a, b)
