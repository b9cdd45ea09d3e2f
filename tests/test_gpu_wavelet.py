"""GPU parity of K5 / K6 (wavelet and superlet transforms) vs golden vectors and the oracle."""
import numpy as np
import pytest

from conftest import load_golden, nerr
from oracle import synth
from oracle import timefreq as otf

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _hm_wavelet(name):
    from syncopy_b200 import hostmath as hm
    return {"morlet6": hm.Morlet(6), "paul4": hm.Paul(4), "dog2": hm.DOG(2)}[name]


@pytest.mark.parametrize("name", ["cwt_morlet6", "cwt_paul4", "cwt_dog2"])
def test_golden_cwt(engine, name):
    from syncopy_b200 import batched
    z, prm = load_golden(name)
    spec = batched.wavelet(z["x"][None], prm["fs"], z["scales"], _hm_wavelet(prm["wavelet"]), polyremoval=None,
                           output="fourier", to_host=True)
    got = spec[0, :, 0].transpose(1, 0, 2)                       # [nScales, N, C]
    assert got.shape == z["spec"].shape
    assert nerr(got, z["spec"]) <= TOL


@pytest.mark.parametrize("name", ["superlet_mult", "superlet_faslt"])
def test_golden_superlet(engine, name):
    from syncopy_b200 import batched
    z, prm = load_golden(name)
    spec = batched.superlet(z["x"][None], prm["fs"], z["scales"], polyremoval=None, output="fourier",
                            to_host=True, **prm["kw"])
    got = spec[0, :, 0].transpose(1, 0, 2)
    assert nerr(got, z["spec"]) <= TOL
    powr = batched.superlet(z["x"][None], prm["fs"], z["scales"], polyremoval=None, output="pow", to_host=True,
                            **prm["kw"])
    want = (z["spec"] * z["spec"].conj()).real
    assert nerr(powr[0, :, 0].transpose(1, 0, 2), want) <= TOL


@pytest.mark.parametrize("n,c,pr,output", [(700, 5, 0, "pow"), (1024, 2, 1, "abs"), (333, 9, None, "fourier")])
def test_wavelet_cf_vs_oracle(engine, n, c, pr, output):
    from syncopy_b200 import compute_functions as cf
    from syncopy_b200 import hostmath as hm
    fs = 400.
    x = synth.white_noise_trial(n, c, 3) + np.linspace(0, 2, n, dtype="f4")[:, None]
    wav_o, wav_g = otf.Morlet(6), hm.Morlet(6)
    foi = np.arange(5., 100., 12.)
    scales = wav_o.scale_from_period(1 / foi)
    kw = dict(toi=None, polyremoval=pr, output=output)
    got = cf.wavelet_cF(x.copy(), slice(None), slice(None),
                        method_kwargs=dict(samplerate=fs, scales=scales, wavelet=wav_g), **kw)
    want = otf.wavelet_cF(x.copy(), slice(None), slice(None),
                          method_kwargs=dict(samplerate=fs, scales=scales, wavelet=wav_o), **kw)
    assert got.shape == want.shape and got.dtype == want.dtype
    assert nerr(got, want) <= TOL
    # dry run
    shp, dt = cf.wavelet_cF(x, slice(None), slice(None), noCompute=True, output=output,
                            method_kwargs=dict(samplerate=fs, scales=scales, wavelet=wav_g))
    assert shp == want.shape and dt == want.dtype


def test_wavelet_cf_time_selection(engine):
    """toi as array: `preselect` cuts the samples, `postselect` picks (possibly repeated) rows."""
    from syncopy_b200 import compute_functions as cf
    from syncopy_b200 import hostmath as hm
    fs, n = 500., 900
    x = synth.white_noise_trial(n, 4, 8)
    scales = otf.Morlet(6).scale_from_period(1 / np.array([10., 40., 90.]))
    pre, post = slice(100, 800), np.array([0, 5, 5, 300, 699])
    toi = np.zeros(post.size)
    got = cf.wavelet_cF(x.copy(), pre, post, toi=toi, polyremoval=0, output="pow",
                        method_kwargs=dict(samplerate=fs, scales=scales, wavelet=hm.Morlet(6)))
    want = otf.wavelet_cF(x.copy(), pre, post, toi=toi, polyremoval=0, output="pow",
                          method_kwargs=dict(samplerate=fs, scales=scales, wavelet=otf.Morlet(6)))
    assert got.shape == want.shape == (5, 1, 3, 4) and nerr(got, want) <= TOL


@pytest.mark.parametrize("adaptive", [False, True])
def test_superlet_cf_vs_oracle(engine, adaptive):
    from syncopy_b200 import compute_functions as cf
    fs, n = 500., 800
    t = np.arange(n) / fs
    x = (np.sin(2 * np.pi * 30 * t) * (t > 0.5) + 0.5 * np.sin(2 * np.pi * 70 * t))[:, None].astype("f4") \
        + 0.1 * synth.white_noise_trial(n, 3, 1)
    foi = np.arange(10., 100., 10.)
    scales = 1 / (2 * np.pi * foi)                                # high -> low scale, as the front-end orders them
    mk = dict(samplerate=fs, scales=scales, order_max=6, order_min=1, c_1=3, adaptive=adaptive)
    got = cf.superlet_cF(x.copy(), slice(None), slice(None), polyremoval=0, output="pow", method_kwargs=dict(mk))
    want = otf.superlet_cF(x.copy(), slice(None), slice(None), polyremoval=0, output="pow", method_kwargs=dict(mk))
    assert got.shape == want.shape and nerr(got, want) <= TOL
    # the 30 Hz packet shows up at the right place: power at 30 Hz after 0.5 s >> before
    k30 = 2
    assert got[int(0.8 * fs):, 0, k30, 0].mean() > 20 * got[:int(0.3 * fs), 0, k30, 0].mean()


def test_cfg5_shape_subset_vs_oracle(engine):
    """BASELINE cfg-5 shape (8192 samples x 64 channels, Morlet w0 = 6, fs = 1000), a subset of the 50 scales so that
    the CPU oracle finishes in seconds; includes the lowest frequency, whose kernel support (9681 taps) exceeds the trial."""
    from syncopy_b200 import batched
    from syncopy_b200 import hostmath as hm
    fs = 1000.
    x = synth.white_noise(2, 8192, 64)
    wav_o, wav_g = otf.Morlet(6), hm.Morlet(6)
    foi = np.array([1., 9., 41., 99.])
    scales = wav_o.scale_from_period(1 / foi)
    got = batched.wavelet(x, fs, scales, wav_g, polyremoval=0, output="pow", to_host=True)
    assert got.shape == (2, 8192, 1, 4, 64)
    for k in range(2):
        want = otf.wavelet_cF(x[k].copy(), slice(None), slice(None), toi=None, polyremoval=0, output="pow",
                              method_kwargs=dict(samplerate=fs, scales=scales, wavelet=wav_o))
        assert nerr(got[k], want) <= TOL


@pytest.mark.parametrize("n,c,output", [(3000, 6, "fourier"), (8192, 3, "pow")])
def test_overlap_save_groups_vs_oracle(engine, n, c, output):
    """scales whose kernels are short run as overlap-save blocks of the trial, the others through the full padded
    length (hostmath.conv_groups): one call mixing both, every scale against the oracle"""
    from syncopy_b200 import compute_functions as cf
    from syncopy_b200 import hostmath as hm
    fs = 1000.
    x = synth.white_noise_trial(n, c, 17) + np.float32(0.3)
    wav_o, wav_g = otf.Morlet(6), hm.Morlet(6)
    foi = np.array([2., 5., 9., 20., 60., 150., 300.])
    scales = wav_o.scale_from_period(1 / foi)
    kw = dict(toi=None, polyremoval=0, output=output)
    got = cf.wavelet_cF(x.copy(), slice(None), slice(None),
                        method_kwargs=dict(samplerate=fs, scales=scales, wavelet=wav_g), **kw)
    want = otf.wavelet_cF(x.copy(), slice(None), slice(None),
                          method_kwargs=dict(samplerate=fs, scales=scales, wavelet=wav_o), **kw)
    assert got.shape == want.shape
    for s in range(foi.size):
        assert nerr(got[:, :, s], want[:, :, s]) <= TOL, foi[s]
    plans = [p for k, p in engine._plans.items() if k[0] == "conv" and k[2] == n]
    assert plans and any(any(g["seg"] for g in p["groups"]) and any(not g["seg"] for g in p["groups"]) for p in plans)


@pytest.mark.parametrize("adaptive", [False, True])
def test_overlap_save_superlet_vs_oracle(engine, adaptive):
    """superlets: a scale's factors (orders) have different supports; a run of scales takes the largest"""
    from syncopy_b200 import compute_functions as cf
    fs, n = 1000., 4000
    x = synth.white_noise_trial(n, 4, 23)
    foi = np.array([160., 80., 45., 20.])
    scales = 1 / (2 * np.pi * foi)
    mk = dict(samplerate=fs, scales=scales, order_max=6, order_min=1, c_1=3, adaptive=adaptive)
    for output in ("fourier", "pow"):
        got = cf.superlet_cF(x.copy(), slice(None), slice(None), polyremoval=0, output=output, method_kwargs=dict(mk))
        want = otf.superlet_cF(x.copy(), slice(None), slice(None), polyremoval=0, output=output, method_kwargs=dict(mk))
        assert got.shape == want.shape
        for k in range(foi.size):
            assert nerr(got[:, :, k], want[:, :, k]) <= TOL, (output, foi[k])
