"""
GPU parity of K1 (tapered FFT) through the C ABI: CUDA path vs the CPU oracle and vs the
committed reference golden vectors.  Tolerance: north_star's 1e-5 relative (normwise, FP32).
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, nerr
from oracle import spectral as osp
from oracle import synth

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _run_cf(x, **kw):
    from syncopy_b200 import compute_functions as cf
    return cf.mtmfft_cF(x, **kw)


@pytest.mark.parametrize("name", ["mtmfft_hann_n1000", "mtmfft_boxcar_odd_n1001",
                                  "mtmfft_dpss_pad2048", "mtmfft_kaiser_ftcompat_n512"])
def test_golden_reference_vectors(engine, name):
    """Full complex spectra against outputs of the real reference."""
    z, prm = load_golden(name)
    kw = dict(prm["kw"])
    mk = dict(samplerate=prm["fs"], nSamples=kw.get("nSamples"), taper=kw.get("taper"),
              taper_opt=kw.get("taper_opt") or {}, demean_taper=kw.get("demean_taper", False),
              ft_compat=kw.get("ft_compat", False))
    nfft = mk["nSamples"] or z["x"].shape[0]
    foi = np.fft.rfftfreq(nfft, 1 / prm["fs"])
    got, meta = _run_cf(z["x"].copy(), foi=foi, keeptapers=True, polyremoval=None, output="fourier",
                        method_kwargs=mk)
    assert got.dtype == np.complex64 and got.shape == (1,) + z["ftr"].shape
    assert nerr(got[0], z["ftr"]) <= TOL
    assert meta["freqs_hash"] == osp.freqs_hash(z["freqs"])


@pytest.mark.parametrize("n,c,nfft", [
    (16, 2, None), (32, 1, None), (64, 5, None), (128, 8, None), (256, 3, 512), (512, 16, None),
    (1024, 32, None), (2048, 9, None), (4096, 24, None), (3000, 4, 4096), (8192, 6, None),
    (16384, 4, None), (10000, 2, 16384),
    (1000, 3, None), (1001, 2, None), (17, 3, None), (7, 2, None), (5000, 2, None), (4095, 2, 6000),
])
def test_sizes_vs_oracle(engine, n, c, nfft):
    """Power-of-two (direct) and arbitrary (Bluestein) lengths, even/odd channel counts."""
    x = np.random.default_rng(n + c).normal(size=(n, c)).astype("f4")
    mk = dict(samplerate=1000., nSamples=nfft, taper="hann", taper_opt={})
    foi = np.fft.rfftfreq(nfft or n, 1e-3)
    got, _ = _run_cf(x.copy(), foi=foi, output="fourier", polyremoval=0, method_kwargs=mk)
    want, _ = osp.mtmfft_cF(x.copy(), foi=foi, output="fourier", polyremoval=0, method_kwargs=mk)
    assert got.shape == want.shape and got.dtype == want.dtype
    assert nerr(got, want) <= TOL


@pytest.mark.parametrize("output", ["pow", "abs", "fourier", "real", "imag", "angle", "absreal", "absimag"])
@pytest.mark.parametrize("keeptapers", [True, False])
def test_outputs_and_taper_mean(engine, output, keeptapers):
    if not keeptapers and output == "angle":
        pytest.skip("averaging phases across tapers is not meaningful (and wraps)")
    x = synth.white_noise_trial(1024, 6, 7) + np.float32(0.5)
    mk = dict(samplerate=512., nSamples=None, taper="dpss", taper_opt={"NW": 3, "Kmax": 5})
    foi = np.fft.rfftfreq(1024, 1 / 512.)
    got, _ = _run_cf(x.copy(), foi=foi, keeptapers=keeptapers, polyremoval=1, output=output, method_kwargs=mk)
    want, _ = osp.mtmfft_cF(x.copy(), foi=foi, keeptapers=keeptapers, polyremoval=1, output=output, method_kwargs=mk)
    assert got.shape == want.shape and got.dtype == want.dtype
    if output == "angle":
        d = np.angle(np.exp(1j * (got.astype(np.float64) - want)))      # compare modulo 2 pi
        # phases of near-zero bins are ill-conditioned; weight by the bin magnitude
        mag, _ = osp.mtmfft_cF(x.copy(), foi=foi, keeptapers=True, polyremoval=1, output="abs", method_kwargs=mk)
        assert np.max(np.abs(d) * mag) / mag.max() <= TOL
    else:
        assert nerr(got, want) <= TOL


def test_foi_gather_unsorted_and_timeaxis(engine):
    x = synth.white_noise_trial(1000, 4, 3)
    mk = dict(samplerate=1000., nSamples=None, taper="hann", taper_opt={})
    foi = np.array([100.2, 40., 40.1, 300., 7.5, 499.9, 0.])
    a, _ = _run_cf(x.T.copy(), foi=foi, timeAxis=1, polyremoval=0, output="pow", method_kwargs=mk)
    b, _ = osp.mtmfft_cF(x.T.copy(), foi=foi, timeAxis=1, polyremoval=0, output="pow", method_kwargs=mk)
    assert a.shape == b.shape == (1, 1, 6, 4)
    assert nerr(a, b) <= TOL


def test_known_amplitudes(engine):
    """Reference known-answer test (tests/backend/test_timefreq.py:351-379) on the GPU path."""
    f1, f2, A1, A2, n, fs = 40, 100, 5, 3, 1000, 1000
    t = np.arange(0, 1, 1 / n)
    sig = (A1 * np.cos(2 * np.pi * f1 * t) + A2 * np.cos(2 * np.pi * f2 * t)).astype("f4")[:, None]
    mk = dict(samplerate=fs, nSamples=None, taper=None, taper_opt={})
    p, _ = _run_cf(sig, foi=np.fft.rfftfreq(n, 1 / fs), output="pow", method_kwargs=mk)
    assert np.allclose([0.5 * A1 ** 2, 0.5 * A2 ** 2], p[0, 0, [f1, f2], 0], rtol=1e-5)


def test_padding_invariant_power(engine):
    """tests/test_specest.py:238-317: the normalisation makes power independent of the padding."""
    n, fs = 1000, 1000
    t = np.arange(n) / fs
    sig = (np.pi * np.cos(2 * np.pi * 50 * t)).astype("f4")[:, None]
    powers = []
    for nfft in (1000, 2000, 4000):
        mk = dict(samplerate=fs, nSamples=nfft, taper="hann", taper_opt={})
        p, _ = _run_cf(sig, foi=np.fft.rfftfreq(nfft, 1 / fs), output="pow", method_kwargs=mk)
        powers.append(p.sum() * (n / nfft))
    assert np.allclose(powers, powers[0], rtol=1e-4)


def test_batched_matches_per_trial_and_trial_indexing(engine):
    """Whole-dataset path == per-trial cF, trial k lands in row block k (bit-exact indexing)."""
    from syncopy_b200 import batched
    trials = synth.white_noise(5, 512, 6)
    for k in range(5):
        trials[k] += np.float32(k)          # make trials distinguishable
    spec, freqs = batched.mtmfft(trials, 1000., taper="hann", polyremoval=None, output="pow", to_host=True)
    assert spec.shape == (5, 1, 257, 6)
    mk = dict(samplerate=1000., nSamples=None, taper="hann", taper_opt={})
    for k in range(5):
        one, _ = _run_cf(trials[k], foi=freqs, output="pow", method_kwargs=mk)
        assert np.array_equal(one[0], spec[k])
    av, _ = batched.mtmfft(trials, 1000., taper="hann", output="pow", keeptrials=False, to_host=True)
    assert nerr(av[0], spec.mean(axis=0)) <= 1e-6


def test_full_size_parseval(engine):
    """BASELINE cfg-2 trial shape (4096 x 256): Parseval as the size-independent property."""
    x = synth.white_noise_trial(4096, 256, 1)
    xd = engine.to_device(x)[None]
    tap = engine.taper_table(None, 4096, 4096)
    spec = engine.mtmfft(xd, tap, 4096, 1.0, output="pow")[0, 0].double()   # raw |X|^2 with boxcar*1
    # boxcar table is scaled by sqrt(P/sum w) = 1 -> plain rfft
    w = torch.ones(2049, dtype=torch.float64, device=spec.device) * 2
    w[0] = w[-1] = 1
    energy_f = (spec * w[:, None]).sum(dim=0) / 4096
    energy_t = (xd[0].double() ** 2).sum(dim=0)
    assert torch.allclose(energy_f, energy_t, rtol=1e-5)


def test_cfg1_reference_frontend_setup(engine):
    """BASELINE cfg-1 = the reference's TestMTMFFT data (tests/test_specest.py:99-150): 8 trials x 32 channels of
    1 s sines (amplitude pi, one frequency per channel) at fs = 1024; every channel peaks at its own frequency
    (:229-233) and the GPU spectrum matches the oracle."""
    from syncopy_b200 import batched
    from oracle import spectral as osp
    fs, n, n_chan, n_trials = 1024, 1024, 32, 8
    rng = np.random.default_rng(42)
    freqs = rng.choice(np.arange(4, 500), size=n_chan, replace=False)
    t = np.arange(n) / fs
    trials = np.stack([np.stack([np.pi * np.sin(2 * np.pi * f * t + rng.uniform(0, 2 * np.pi)) for f in freqs], axis=1)
                       for _ in range(n_trials)]).astype("f4")
    spec, f_axis = batched.mtmfft(trials, fs, taper="hann", polyremoval=0, output="pow", keeptapers=False,
                                  to_host=True)
    assert spec.shape == (n_trials, 1, n // 2 + 1, n_chan)
    for k in range(n_trials):
        assert np.array_equal(f_axis[spec[k, 0].argmax(axis=0)], freqs.astype(float))
    mk = dict(samplerate=fs, nSamples=None, taper="hann", taper_opt={})
    want, _ = osp.mtmfft_cF(trials[3].copy(), foi=f_axis, polyremoval=0, output="pow", keeptapers=False,
                            method_kwargs=mk)
    assert nerr(spec[3:4], want) <= 1e-5
