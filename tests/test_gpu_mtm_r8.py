"""
GPU parity of the 512-point tapered-FFT kernel (csrc/mtm_r8.cu: radix-8 passes, frame resident in registers across
tapers, taper sums in registers) against the oracle: detrending modes, zero-padded windows, several tapers with and
without the taper mean, every output conversion, sliding frames with zero-padded boundaries (mtmconvol of BASELINE
cfg-3).  Reference: syncopy/specest/mtmfft.py:16-129, stft.py:95-157, mtmconvol.py:17-152, compRoutines.py:169-189,
410-413.  Tolerance 1e-5 normwise.
"""
import numpy as np
import pytest

from conftest import nerr
from oracle import spectral as osp
from oracle import synth

pytestmark = pytest.mark.gpu
TOL = 1e-5
N = 512


def _cf(x, **kw):
    from syncopy_b200 import compute_functions as cf
    return cf.mtmfft_cF(x, **kw)


@pytest.mark.parametrize("output", ["pow", "abs", "fourier", "real", "imag", "absreal", "absimag"])
@pytest.mark.parametrize("keeptapers", [True, False])
def test_outputs_and_taper_mean(engine, output, keeptapers):
    x = synth.white_noise_trial(N, 24, 7) + np.float32(0.05)
    mk = dict(samplerate=512., nSamples=None, taper="dpss", taper_opt={"NW": 3, "Kmax": 5}, demean_taper=True)
    foi = np.fft.rfftfreq(N, 1 / 512.)
    got, _ = _cf(x.copy(), foi=foi, keeptapers=keeptapers, polyremoval=1, output=output, method_kwargs=mk)
    want, _ = osp.mtmfft_cF(x.copy(), foi=foi, keeptapers=keeptapers, polyremoval=1, output=output, method_kwargs=mk)
    assert got.shape == want.shape and got.dtype == want.dtype
    assert nerr(got, want) <= TOL


@pytest.mark.parametrize("n_sig,pr", [(512, None), (511, 0), (300, 1), (9, 0)])
def test_windows_and_detrend(engine, n_sig, pr):
    x = synth.white_noise_trial(n_sig, 8, n_sig) + np.float32(0.02)
    mk = dict(samplerate=500., nSamples=N, taper="hann", taper_opt={})
    foi = np.fft.rfftfreq(N, 1 / 500.)
    got, _ = _cf(x.copy(), foi=foi, output="fourier", polyremoval=pr, method_kwargs=mk)
    want, _ = osp.mtmfft_cF(x.copy(), foi=foi, output="fourier", polyremoval=pr, method_kwargs=mk)
    assert nerr(got, want) <= TOL


@pytest.mark.parametrize("pr,keeptapers,output", [(0, False, "pow"), (1, True, "fourier"), (None, False, "abs")])
def test_sliding_frames(engine, pr, keeptapers, output):
    """mtmconvol with nperseg = 512, hop 256 (cfg-3 geometry): boundary frames are zero padded and de-meaned with their
    zeros (stft.py:101-132); odd tile counts leave one tile of the last block idle"""
    from syncopy_b200 import batched
    x = synth.white_noise_trial(3000, 24, 5) + np.float32(0.02)
    spec, freqs = batched.mtmconvol(x[None], 1024., N, 256, taper="dpss", taper_opt={"NW": 4, "Kmax": 7},
                                    polyremoval=pr, output=output, keeptapers=keeptapers, to_host=True)
    det = {0: "constant", 1: "linear", None: False}[pr]
    ftr, _ = osp.mtmconvol(x.copy(), 1024., N, 256, "dpss", {"NW": 4, "Kmax": 7}, "zeros", True, det)
    want = osp.convert_output(ftr, output)
    if not keeptapers:
        want = want.mean(axis=1, keepdims=True)
    assert spec.shape == (1,) + want.shape
    assert nerr(spec[0], want) <= TOL


def test_planar_and_batch_bitwise(engine):
    import torch
    from syncopy_b200 import hostmath as hm
    x = torch.from_numpy(synth.white_noise(5, N, 16)).to(engine.tdev)
    tapers = engine.taper_table("dpss", N, N, {"NW": 2, "Kmax": 3})
    scale = hm.mtmfft_scale(N, N)
    planes = engine.mtmfft(x, tapers, N, scale, polyremoval=0, output="fourier_planar", keeptapers=True, freq_major=True)
    inter = engine.mtmfft(x, tapers, N, scale, polyremoval=0, output="fourier", keeptapers=True, freq_major=True)
    assert torch.equal(torch.view_as_real(inter)[..., 0], planes[:, :, 0, :])
    assert torch.equal(torch.view_as_real(inter)[..., 1], planes[:, :, 1, :])
    one = engine.mtmfft(x[3:4], tapers, N, scale, polyremoval=0, output="fourier", keeptapers=True, freq_major=True)
    assert torch.equal(one, inter[:, 9:12])
