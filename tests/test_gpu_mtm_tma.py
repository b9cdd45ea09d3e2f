"""
GPU parity of the pipelined tapered-FFT kernel (csrc/mtm_tma.cu: N = 4096, 8-channel tiles, TMA tile loads, packed
FP32 butterflies) against the oracle -- every option the kernel folds in: detrending modes, zero padding
(window shorter than the FFT), several tapers, post-taper de-meaning, every output conversion, the planar layout
handed to the cross-spectral kernel, trial strides, frames (mtmconvol with nperseg = 4096 incl. zero-padded
boundary frames) and more tiles than SMs (persistent loop, in-place refills).
Reference: syncopy/specest/mtmfft.py:16-129, compRoutines.py:169-189, stft.py:95-157.  Tolerance 1e-5 normwise.
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


def _cf(x, **kw):
    from syncopy_b200 import compute_functions as cf
    return cf.mtmfft_cF(x, **kw)


def _exact_detrended(x, n_sig, polyremoval):
    """float64 detrend (scipy.signal.detrend semantics), rounded once to float32"""
    x64 = x.astype(np.float64)
    if polyremoval == 0:
        x64 = x64 - x64.mean(axis=0, keepdims=True)
    elif polyremoval == 1:
        t = np.arange(n_sig, dtype=np.float64)
        A = np.stack([t, np.ones(n_sig)], axis=1)
        coef, *_ = np.linalg.lstsq(A, x64, rcond=None)
        x64 = x64 - A @ coef
    return x64.astype("f4")


@pytest.mark.parametrize("n_chan", [8, 24, 256])
@pytest.mark.parametrize("polyremoval", [None, 0, 1])
def test_detrend_modes(engine, n_chan, polyremoval):
    x = synth.white_noise_trial(N, n_chan, 3 + n_chan)
    x += np.linspace(-0.04, 0.05, n_chan, dtype="f4")[None, :]
    x += np.linspace(-0.03, 0.02, N, dtype="f4")[:, None] * np.linspace(0.5, 2.0, n_chan, dtype="f4")[None, :]
    mk = dict(samplerate=1000., nSamples=None, taper="hann", taper_opt={})
    foi = np.fft.rfftfreq(N, 1e-3)
    got, _ = _cf(x.copy(), foi=foi, output="fourier", polyremoval=polyremoval, method_kwargs=mk)
    want, _ = osp.mtmfft_cF(x.copy(), foi=foi, output="fourier", polyremoval=polyremoval, method_kwargs=mk)
    assert got.shape == want.shape == (1, 1, N // 2 + 1, n_chan)
    assert nerr(got, want) <= TOL


@pytest.mark.parametrize("polyremoval", [0, 1])
def test_offsets_vs_float64_detrend(engine, polyremoval):
    """Offsets / trends of several standard deviations: the reference detrends in float32 with sequential column sums
    (scipy.signal.detrend on float32 input, compRoutines.py:169-172), which alone costs it 1e-4 of the spectrum's
    maximum at 4096 samples; the kernel reduces pairwise.  Bar: within tolerance of the same chain detrended in
    float64, and never further from it than the reference is."""
    n_chan = 32
    x = synth.white_noise_trial(N, n_chan, 259) + np.float32(2.5)
    x += np.linspace(-1.0, 1.5, N, dtype="f4")[:, None]
    mk = dict(samplerate=1000., nSamples=None, taper="hann", taper_opt={})
    foi = np.fft.rfftfreq(N, 1e-3)
    got, _ = _cf(x.copy(), foi=foi, output="fourier", polyremoval=polyremoval, method_kwargs=mk)
    ref, _ = osp.mtmfft_cF(x.copy(), foi=foi, output="fourier", polyremoval=polyremoval, method_kwargs=mk)
    exact, _ = osp.mtmfft_cF(_exact_detrended(x, N, polyremoval), foi=foi, output="fourier", polyremoval=None,
                             method_kwargs=mk)
    e_gpu, e_ref = nerr(got, exact), nerr(ref, exact)
    assert e_gpu <= 2 * TOL and e_gpu <= max(e_ref, TOL), (e_gpu, e_ref)


@pytest.mark.parametrize("n_sig", [4095, 3000, 33, 257])
@pytest.mark.parametrize("polyremoval", [0, 1])
def test_zero_padded_window(engine, n_sig, polyremoval):
    """trial shorter than the FFT (pad='nextpow2' / nSamples): rows past the window are zero, the detrending sums
    and the taper only see the window"""
    x = synth.white_noise_trial(n_sig, 16, n_sig) + np.float32(0.03)
    mk = dict(samplerate=500., nSamples=N, taper="hann", taper_opt={})
    foi = np.fft.rfftfreq(N, 1 / 500.)
    got, _ = _cf(x.copy(), foi=foi, output="fourier", polyremoval=polyremoval, method_kwargs=mk)
    want, _ = osp.mtmfft_cF(x.copy(), foi=foi, output="fourier", polyremoval=polyremoval, method_kwargs=mk)
    assert got.shape == want.shape
    assert nerr(got, want) <= TOL


@pytest.mark.parametrize("output", ["pow", "abs", "fourier", "real", "imag", "angle", "absreal", "absimag"])
def test_output_kinds_multitaper(engine, output):
    x = synth.white_noise_trial(N, 16, 21) + np.float32(0.03)
    mk = dict(samplerate=1000., nSamples=None, taper="dpss", taper_opt={"NW": 3, "Kmax": 5}, demean_taper=True)
    foi = np.fft.rfftfreq(N, 1e-3)
    got, _ = _cf(x.copy(), foi=foi, keeptapers=True, polyremoval=0, output=output, method_kwargs=mk)
    want, _ = osp.mtmfft_cF(x.copy(), foi=foi, keeptapers=True, polyremoval=0, output=output, method_kwargs=mk)
    assert got.shape == want.shape == (1, 5, N // 2 + 1, 16) and got.dtype == want.dtype
    if output == "angle":
        d = np.angle(np.exp(1j * (got.astype(np.float64) - want)))
        mag, _ = osp.mtmfft_cF(x.copy(), foi=foi, keeptapers=True, polyremoval=0, output="abs", method_kwargs=mk)
        assert np.max(np.abs(d) * mag) / mag.max() <= TOL
    else:
        assert nerr(got, want) <= TOL


def test_planar_layout_and_trial_stride(engine):
    """out_kind 8 ([f][row][re|im][c], what K2 consumes) from a strided view of a larger trial buffer, more tiles
    (23 trials x 32 channel tiles = 736) than SMs, three tapers"""
    from syncopy_b200 import hostmath as hm
    rng = np.random.default_rng(8)
    big = torch.from_numpy(rng.normal(size=(23, N + 40, 256)).astype("f4")).to(engine.tdev)
    x = big[:, 8:8 + N]                                      # trial stride (N + 40) * 256
    tapers = engine.taper_table("dpss", N, N, {"NW": 2, "Kmax": 3})
    scale = hm.mtmfft_scale(N, N)
    planes = engine.mtmfft(x, tapers, N, scale, polyremoval=0, output="fourier_planar", keeptapers=True,
                           freq_major=True)
    assert planes.shape == (N // 2 + 1, 23 * 3, 2, 256)
    got = torch.complex(planes[:, :, 0, :], planes[:, :, 1, :]).cpu().numpy()      # [f, trial*3 + taper, c]
    xh = x.cpu().numpy()
    mk = dict(samplerate=1000., nSamples=None, taper="dpss", taper_opt={"NW": 2, "Kmax": 3})
    foi = np.fft.rfftfreq(N, 1e-3)
    for t in (0, 7, 22):
        want, _ = osp.mtmfft_cF(xh[t].copy(), foi=foi, keeptapers=True, polyremoval=0, output="fourier",
                                method_kwargs=mk)            # [1, 3, nF, C]
        for k in range(3):
            assert nerr(got[:, t * 3 + k, :], want[0, k]) <= TOL
    # interleaved complex output of the same launch shape must agree bit for bit with the planar one
    inter = engine.mtmfft(x, tapers, N, scale, polyremoval=0, output="fourier", keeptapers=True, freq_major=True)
    assert torch.equal(torch.view_as_real(inter)[..., 0], planes[:, :, 0, :])
    assert torch.equal(torch.view_as_real(inter)[..., 1], planes[:, :, 1, :])


def test_batched_equals_single_trial_bitwise(engine):
    """persistent-loop bookkeeping: trial k of a batch == the same trial run alone, bit for bit"""
    from syncopy_b200 import batched
    trials = synth.white_noise(7, N, 40)
    for k in range(7):
        trials[k] += np.float32(k)
    spec, freqs = batched.mtmfft(trials, 1000., taper="hann", polyremoval=0, output="pow", to_host=True)
    assert spec.shape == (7, 1, N // 2 + 1, 40)
    for k in (0, 3, 6):
        one, _ = batched.mtmfft(trials[k:k + 1], 1000., taper="hann", polyremoval=0, output="pow", to_host=True)
        assert np.array_equal(one[0], spec[k])


@pytest.mark.parametrize("boundary,pr", [("zeros", 0), ("zeros", 1), ("zeros", None)])
def test_frames_nperseg_4096(engine, boundary, pr):
    """mtmconvol with nperseg = 4096: frames start at negative sample indices / run past the trial end (zero fill
    by TMA), per-segment detrending includes those zeros (stft.py:101-132)"""
    from syncopy_b200 import batched
    x = synth.white_noise_trial(10000, 8, 5) + np.float32(0.03)
    spec, freqs = batched.mtmconvol(x[None], 1000., N, N - 1500, taper="hann", boundary=boundary, padded=True,
                                    polyremoval=pr, output="fourier", to_host=True)
    det = {0: "constant", 1: "linear", None: False}[pr]
    ftr, _ = osp.mtmconvol(x.copy(), 1000., N, N - 1500, "hann", {}, boundary, True, det)
    assert spec.shape == (1,) + ftr.shape
    assert nerr(spec[0], ftr) <= TOL
