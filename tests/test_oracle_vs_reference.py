"""Oracle == live reference (by-path import) on fresh seeded inputs.  Only runs where
/root/reference exists (the build container); skipped on the GPU box."""
import numpy as np
import pytest

from conftest import nerr
from oracle import connectivity as oc
from oracle import ref_loader
from oracle import spectral as osp
from oracle import timefreq as otf

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


@pytest.mark.parametrize("n,c,nfft,taper,opt,demean", [
    (1024, 8, None, "hann", None, False),
    (777, 3, 1000, "dpss", {"NW": 4, "Kmax": 7}, True),
    (256, 5, 512, None, None, False),
    (300, 2, None, "hamming", {}, True),
])
def test_mtmfft_live(ref, n, c, nfft, taper, opt, demean):
    x = np.random.default_rng(n).normal(size=(n, c)).astype("f4")
    want, fr = ref.mtmfft.mtmfft(x.copy(), 500., nfft, taper, opt, demean)
    got, fr2 = osp.mtmfft(x.copy(), 500., nfft, taper, opt, demean)
    assert nerr(got, want) <= 1e-12 and np.array_equal(fr, fr2)


def test_csd_live(ref):
    x = np.random.default_rng(5).normal(size=(400, 7)).astype("f4")
    want, _ = ref.csd.csd(x.copy(), 1000., 512, "dpss", {"NW": 3, "Kmax": 5}, True)
    got, _ = oc.csd(x.copy(), 1000., 512, "dpss", {"NW": 3, "Kmax": 5}, True)
    assert nerr(got, want) <= 1e-12
    with pytest.raises(Exception):
        oc.csd(x, 1000., taper="hann", norm=True)


@pytest.mark.parametrize("bdry,padded,det,nover", [("zeros", True, "constant", 96), (None, False, False, 127),
                                                  ("zeros", True, "linear", 0)])
def test_mtmconvol_live(ref, bdry, padded, det, nover):
    x = np.random.default_rng(9).normal(size=(900, 3)).astype("f4")
    kw = dict(nperseg=128, noverlap=nover, taper="dpss", taper_opt={"NW": 2, "Kmax": 3},
              boundary=bdry, padded=padded, detrend=det)
    want, _ = ref.mtmconvol.mtmconvol(x.copy(), 400., **{**kw, "taper_opt": dict(kw["taper_opt"])})
    got, _ = osp.mtmconvol(x.copy(), 400., **kw)
    assert nerr(got, want) <= 1e-12


def test_best_match_live(ref):
    rng = np.random.default_rng(3)
    src = np.fft.rfftfreq(501, 1 / 250.)
    for _ in range(20):
        sel = rng.uniform(-5, 140, size=rng.integers(1, 30))
        for squash in (False, True):
            assert np.array_equal(ref.tools.best_match(src, sel, squash_duplicates=squash)[1],
                                  osp.best_match(src, sel, squash_duplicates=squash)[1])


def test_superlet_cwt_live(ref):
    x = np.random.default_rng(11).normal(size=(500, 2)).astype("f4")
    scales = ref.superlet.scale_from_period(1 / np.linspace(20, 120, 6))
    for adaptive in (False, True):
        want = ref.superlet.superlet(x.copy(), 500., scales, 4, 1, 3, adaptive)
        got = otf.superlet(x.copy(), 500., scales, 4, 1, 3, adaptive)
        assert nerr(got, want) <= 1e-6
    w = ref.wavelets_mod.Morlet(6)
    sc = w.scale_from_period(1 / np.array([5., 20., 80.]))
    assert nerr(otf.wavelet(x.copy(), 500., sc, otf.Morlet(6)), ref.wavelet.wavelet(x.copy(), 500., sc, w)) <= 1e-7


def _mvar_like_csd(n_chan, n_freq, seed):
    """Well-conditioned Hermitian positive-definite spectral matrices (one-sided), real at DC and Nyquist."""
    rng = np.random.default_rng(seed)
    a1 = 0.4 * np.eye(n_chan) + 0.2 / np.sqrt(n_chan) * rng.normal(size=(n_chan, n_chan))
    q = rng.normal(size=(n_chan, n_chan))
    sigma = q @ q.T / n_chan + np.eye(n_chan)
    om = np.pi * np.arange(n_freq) / (n_freq - 1)
    Hf = np.linalg.inv(np.eye(n_chan)[None] - a1[None] * np.exp(-1j * om)[:, None, None])
    S = Hf @ sigma[None] @ Hf.conj().transpose(0, 2, 1)
    return 0.5 * (S + S.conj().transpose(0, 2, 1))


@pytest.mark.parametrize("n_chan,n_freq", [(2, 33), (5, 26), (9, 65)])
def test_wilson_granger_live(ref, n_chan, n_freq):
    S = _mvar_like_csd(n_chan, n_freq, n_chan)
    H, Sig, conv, err = ref.wilson_sf.wilson_sf(S.copy(), nIter=80, rtol=1e-9)
    H2, Sig2, conv2, err2 = oc.wilson_sf(S.copy(), nIter=80, rtol=1e-9)
    assert conv == conv2 and nerr(H2, H) <= 1e-12 and nerr(Sig2, Sig) <= 1e-12
    assert nerr(oc.granger(S, H2, Sig2), ref.granger.granger(S, H, Sig)) <= 1e-12
    # iteration cap: not converged is a normal return, identical partial result
    Hc, Sc, cc, ec = ref.wilson_sf.wilson_sf(S.copy(), nIter=2, rtol=1e-14)
    Hc2, Sc2, cc2, ec2 = oc.wilson_sf(S.copy(), nIter=2, rtol=1e-14)
    assert (cc, cc2) == (False, False) and nerr(Hc2, Hc) <= 1e-12 and abs(ec - ec2) <= 1e-12 * ec


@pytest.mark.parametrize("cond_max,scale", [(1e3, 1.0), (5.0, 1e-2), (3.0, 1e-4), (10.0, 1e3)])
def test_regularize_live(ref, cond_max, scale):
    S = (_mvar_like_csd(6, 20, 6) * scale).astype(np.complex64)
    want = ref.wilson_sf.regularize_csd(S.copy(), cond_max=cond_max, eps_max=1e-1)
    got = oc.regularize_csd(S.copy(), cond_max=cond_max, eps_max=1e-1)
    assert got[1] == want[1] and float(got[2]) == float(want[2])
    assert np.array_equal(np.asarray(got[0]), np.asarray(want[0]))


def test_normalize_csd_live(ref):
    S = _mvar_like_csd(4, 17, 2).astype(np.complex64)
    for output in ("abs", "pow", "fourier", "real", "imag", "angle", "absreal", "absimag"):
        want = ref.csd.normalize_csd(S.copy(), output)
        got = oc.normalize_csd(S.copy(), output)
        assert got.dtype == want.dtype and nerr(got, want) <= 1e-12


# ---------------------------------------------------------------------------------------------------------
# "next" rows: functions whose modules need h5py -- the function bodies are compiled from the reference
# file itself (ref_loader.extract_function) and run side by side with the restatement
# ---------------------------------------------------------------------------------------------------------
def _ref_st_function(ref, name, extra=None):
    import scipy.signal
    env = dict(np=np, detrend=scipy.signal.detrend, fftconvolve=scipy.signal.fftconvolve,
               spectralDTypes=ref.const_def.spectralDTypes, spectralConversions=ref.const_def.spectralConversions)
    env.update(extra or {})
    return ref_loader.extract_function("connectivity/ST_compRoutines.py", name, env)


@pytest.mark.parametrize("n,pr,norm", [(100, 0, False), (101, 1, True), (64, None, False), (33, 0, True)])
def test_cross_covariance_live(ref, n, pr, norm):
    from oracle import statistics as ost
    fn = _ref_st_function(ref, "cross_covariance_cF")
    x = np.random.default_rng(n).normal(size=(n, 4)).astype("f4") + np.float32(0.3)
    want, lags = fn(x.copy(), samplerate=200., polyremoval=pr, norm=norm, fullOutput=True)
    got, lags2 = ost.cross_covariance_cF(x.copy(), samplerate=200., polyremoval=pr, norm=norm, fullOutput=True)
    assert got.shape == want.shape and np.array_equal(lags, lags2)
    assert nerr(got, want) <= 1e-12
    assert fn(x, noCompute=True)[0] == ost.cross_covariance_cF(x, noCompute=True)[0]


def test_ppc_column_live(ref):
    """ppc_column_cF reads the second trial from HDF5: hand it a stand-in for `h5py.File`."""
    from oracle import statistics as ost
    rng = np.random.default_rng(2)
    a = (rng.normal(size=(1, 6, 3, 3)) + 1j * rng.normal(size=(1, 6, 3, 3))).astype(np.complex64)
    b = (rng.normal(size=(1, 6, 3, 3)) + 1j * rng.normal(size=(1, 6, 3, 3))).astype(np.complex64)

    class _File:
        def __init__(self, path, mode):
            self.path = path

        def __enter__(self):
            return {"data": {"second": b}}

        def __exit__(self, *exc):
            return False

    import types
    fake_h5py = types.SimpleNamespace(File=_File)
    fn = _ref_st_function(ref, "ppc_column_cF", dict(h5py=fake_h5py))
    want = fn(a, trl2_idx="second", hdf5_path="unused")
    assert np.array_equal(want, ost.ppc_column_cF(a, b))


class _FakeData:
    """Just enough of a Syncopy data object (trials stacked along dim 0) for jackknifing.py's function bodies."""
    _stackingDim = 0
    _filename = "unused"

    def __init__(self, data=None, samplerate=1.0, dimord=None, n_trials=None):
        self.samplerate, self.dimord, self.selection, self.cfg = samplerate, dimord, None, {}
        self._data = None if data is None else np.asarray(data)
        self._n = n_trials

    # -- the attributes the reference touches
    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, value):
        self._data = np.asarray(value) if not isinstance(value, np.ndarray) else value

    def _reopen(self):
        pass

    @property
    def n_trials(self):
        return self._n if self._n is not None else 1

    @property
    def sampleinfo(self):
        step = self._data.shape[0] // self.n_trials
        return np.array([[k * step, (k + 1) * step] for k in range(self.n_trials)])

    @property
    def trials(self):
        step = self._data.shape[0] // self.n_trials
        return [self._data[k * step:(k + 1) * step] for k in range(self.n_trials)]

    def selectdata(self, inplace=True):
        self.selection = type("Sel", (), {"trial_ids": list(range(self.n_trials))})()

    def __sub__(self, other):
        return _FakeData(self._data - other._data, self.samplerate, self.dimord)

    def __rmul__(self, fac):
        return _FakeData(fac * self._data, self.samplerate, self.dimord)

    trialdefinition = None


def _jackknife_functions(ref):
    from oracle import statistics as ost

    class _H5File:
        def __init__(self, name, mode="r"):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

        def create_dataset(self, name, shape, dtype):
            return np.zeros(shape, dtype=dtype)

    import types

    def spy_mean(obj, dim):
        assert dim == "trials"
        return _FakeData(ost.trial_mean(obj.trials), obj.samplerate, obj.dimord)      # summary_stats.py:408-428

    env = dict(np=np, h5py=types.SimpleNamespace(File=_H5File), spy=types.SimpleNamespace(mean=spy_mean),
               propagate_properties=lambda *a, **k: None, SPYValueError=ValueError, SPYError=RuntimeError)
    return (ref_loader.extract_function("statistics/jackknifing.py", "trial_avg_replicates", env),
            ref_loader.extract_function("statistics/jackknifing.py", "bias_var", env))


@pytest.mark.parametrize("dtype", [np.complex64, np.float32])
def test_jackknife_live(ref, dtype):
    """trial_avg_replicates / bias_var: the reference's own function bodies on stand-in data objects."""
    from oracle import statistics as ost
    ref_replicates, ref_bias_var = _jackknife_functions(ref)
    rng = np.random.default_rng(4)
    T = 7
    x = rng.normal(size=(T, 5, 3, 3))
    if dtype is np.complex64:
        x = x + 1j * rng.normal(size=x.shape)
    x = x.astype(dtype)
    ens = _FakeData(x.reshape(T * 1, 5, 3, 3), n_trials=T)
    reps = ref_replicates(ens)
    mine = ost.trial_avg_replicates(list(x[:, None]))
    assert np.array_equal(reps.data.reshape(mine.shape), mine)
    direct = _FakeData(ost.trial_mean(list(x[:, None])), n_trials=1)
    rep_obj = _FakeData(mine.reshape(T, 5, 3, 3), n_trials=T)
    bias, var = ref_bias_var(direct, rep_obj)
    b2, v2 = ost.bias_var(direct.data, list(mine))
    assert np.array_equal(bias.data, b2) and np.array_equal(var.data, v2)


# ---- preprocessing: firws.py / resampling.py load by path, the cF bodies are compiled in place ----------
def _ref_preproc_cf(ref, name):
    import logging
    import platform
    import scipy.signal
    env = dict(np=np, sci=scipy.signal, logging=logging, platform=platform,
               design_wsinc=ref.firws.design_wsinc, apply_fir=ref.firws.apply_fir, minphaserceps=ref.firws.minphaserceps,
               downsample=ref.resampling.downsample, resample=ref.resampling.resample,
               spectralDTypes=ref.const_def.spectralDTypes, spectralConversions=ref.const_def.spectralConversions)
    return ref_loader.extract_function("preproc/compRoutines.py", name, env)


def _noise(n, c, seed):
    return np.random.default_rng(seed).normal(size=(n, c)).astype("f4") + np.float32(0.2)


@pytest.mark.parametrize("ft,freq,order,direction,pr", [
    ("lp", 30., None, "onepass", None), ("hp", 10., 100, "twopass", 0), ("bp", np.array([10., 40.]), 201, "onepass", 1),
    ("bs", np.array([45., 55.]), 300, "onepass-minphase", None)])
def test_sinc_filtering_live(ref, ft, freq, order, direction, pr):
    from oracle import preproc as opp
    fn = _ref_preproc_cf(ref, "sinc_filtering_cF")
    x = _noise(500, 3, 1)
    kw = dict(samplerate=200., filter_type=ft, freq=freq, order=order, direction=direction, polyremoval=pr)
    want, meta = fn(x.copy(), **kw)
    got, meta2 = opp.sinc_filtering_cF(x.copy(), **kw)
    assert nerr(got, want) <= 1e-12 and bool(meta["has_nan"]) == bool(meta2["has_nan"])


@pytest.mark.parametrize("ft,freq,order,direction", [("lp", 30., 6, "twopass"), ("hp", 5., 4, "onepass"),
                                                      ("bp", [10., 40.], 3, "twopass"), ("bs", [45., 55.], 2, "onepass")])
def test_but_filtering_live(ref, ft, freq, order, direction):
    from oracle import preproc as opp
    fn = _ref_preproc_cf(ref, "but_filtering_cF")
    x = _noise(400, 4, 2)
    kw = dict(samplerate=200., filter_type=ft, freq=freq, order=order, direction=direction, polyremoval=0)
    want, _ = fn(x.copy(), **kw)
    got, _ = opp.but_filtering_cF(x.copy(), **kw)
    assert nerr(got, want) <= 1e-12


def test_hilbert_resample_misc_live(ref):
    from oracle import preproc as opp
    x = _noise(333, 3, 3)
    for output in ("abs", "complex", "angle", "real", "imag"):
        assert nerr(opp.hilbert_cF(x.copy(), output=output), _ref_preproc_cf(ref, "hilbert_cF")(x.copy(), output=output)) <= 1e-12
    for kw in (dict(samplerate=1000., new_samplerate=250.), dict(samplerate=500., new_samplerate=333., order=100),
               dict(samplerate=200., new_samplerate=300., lpfreq=60.)):
        want = _ref_preproc_cf(ref, "resample_cF")(x.copy(), **kw)
        got = opp.resample_cF(x.copy(), **kw)
        assert got.shape == want.shape and nerr(got, want) <= 1e-12
        assert opp.resample_cF(x, noCompute=True, **kw) == _ref_preproc_cf(ref, "resample_cF")(x, noCompute=True, **kw)
    assert np.array_equal(opp.rectify_cF(x), _ref_preproc_cf(ref, "rectify_cF")(x))
    assert nerr(opp.standardize_cF(x.copy(), polyremoval=1), _ref_preproc_cf(ref, "standardize_cF")(x.copy(), polyremoval=1)) <= 1e-12
    a, _ = opp.detrending_cF(x.copy(), polyremoval=1)
    b, _ = _ref_preproc_cf(ref, "detrending_cF")(x.copy(), polyremoval=1)
    assert nerr(a, b) <= 1e-12
