"""
GPU parity of the preprocessing compute functions (syncopy/preproc/compRoutines.py) against the
oracle restatements pinned to the reference (tests/test_oracle_vs_reference.py).  Tolerance 1e-5 normwise (the
reference computes in float64 and the results are stored as float32).
"""
import numpy as np
import pytest

from conftest import nerr
from oracle import preproc as opp

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _signal(n, c, seed, fs=500.):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    x = rng.normal(size=(n, c)).astype("f4")
    x += (2.0 * np.sin(2 * np.pi * 12.0 * t) + np.sin(2 * np.pi * 60.0 * t))[:, None].astype("f4")
    return x + np.float32(0.1)


@pytest.mark.parametrize("ft,freq,order,direction,pr", [
    ("lp", 30., None, "onepass", None), ("hp", 10., 100, "twopass", 0), ("bp", np.array([10., 40.]), 201, "onepass", 1),
    ("bs", np.array([55., 65.]), 300, "onepass-minphase", None), ("lp", 100., 1000, "onepass", 0)])
@pytest.mark.parametrize("n,c", [(1000, 5), (4096, 32)])
def test_sinc_filtering_cf(engine, ft, freq, order, direction, pr, n, c):
    from syncopy_b200 import preproc as pp
    x = _signal(n, c, 1)
    kw = dict(samplerate=500., filter_type=ft, freq=freq, order=order, direction=direction, polyremoval=pr)
    got, meta = pp.sinc_filtering_cF(x.copy(), **kw)
    want, meta0 = opp.sinc_filtering_cF(x.copy(), **kw)
    assert got.shape == want.shape and got.dtype == np.float32
    assert nerr(got, want) <= TOL and not bool(meta["has_nan"])
    assert pp.sinc_filtering_cF(x, noCompute=True) == (x.shape, np.float32)


@pytest.mark.parametrize("ft,freq,order,direction,pr", [("lp", 30., 6, "twopass", None), ("hp", 5., 4, "onepass", 0),
                                                         ("bp", [10., 40.], 3, "twopass", 1), ("bs", [55., 65.], 2, "onepass", None),
                                                         ("lp", 80., 8, "twopass", 0)])
@pytest.mark.parametrize("n,c", [(700, 3), (4096, 130)])
def test_but_filtering_cf(engine, ft, freq, order, direction, pr, n, c):
    from syncopy_b200 import preproc as pp
    x = _signal(n, c, 2)
    kw = dict(samplerate=500., filter_type=ft, freq=freq, order=order, direction=direction, polyremoval=pr)
    got, meta = pp.but_filtering_cF(x.copy(), **kw)
    want, _ = opp.but_filtering_cF(x.copy(), **kw)
    assert got.shape == want.shape and got.dtype == np.float32
    assert nerr(got, want) <= TOL


@pytest.mark.parametrize("output", ["abs", "complex", "angle", "real", "imag", "pow"])
@pytest.mark.parametrize("n", [512, 777, 4096])
def test_hilbert_cf(engine, output, n):
    from syncopy_b200 import preproc as pp
    x = _signal(n, 6, 3)
    got = pp.hilbert_cF(x.copy(), output=output)
    want = opp.hilbert_cF(x.copy(), output=output)
    assert got.shape == want.shape and got.dtype == want.dtype
    if output == "angle":
        mag = opp.hilbert_cF(x.copy(), output="abs")
        d = np.angle(np.exp(1j * (got.astype(np.float64) - want)))
        assert np.max(np.abs(d) * mag) / mag.max() <= TOL
    else:
        assert nerr(got, want) <= TOL


@pytest.mark.parametrize("kw", [dict(samplerate=1000., new_samplerate=250.), dict(samplerate=500., new_samplerate=333., order=100),
                                dict(samplerate=200., new_samplerate=300., lpfreq=60.), dict(samplerate=1000., new_samplerate=500., order=2000)])
@pytest.mark.parametrize("n,c", [(600, 4), (4096, 64)])
def test_resample_cf(engine, kw, n, c):
    from syncopy_b200 import preproc as pp
    x = _signal(n, c, 4, fs=kw["samplerate"])
    got = pp.resample_cF(x.copy(), **kw)
    want = opp.resample_cF(x.copy(), **kw)
    assert got.shape == want.shape and got.dtype == np.float32
    assert nerr(got, want) <= TOL
    assert pp.resample_cF(x, noCompute=True, **kw) == opp.resample_cF(x, noCompute=True, **kw)


def test_small_cfs(engine):
    from syncopy_b200 import preproc as pp
    x = _signal(1001, 7, 5)
    assert np.array_equal(pp.rectify_cF(x.copy()), opp.rectify_cF(x.copy()))
    for pr in (0, 1):
        a, ma = pp.detrending_cF(x.copy(), polyremoval=pr)
        b, _ = opp.detrending_cF(x.copy(), polyremoval=pr)
        assert nerr(a, b) <= TOL and not bool(ma["has_nan"])
    for pr in (None, 0, 1):
        assert nerr(pp.standardize_cF(x.copy(), polyremoval=pr), opp.standardize_cF(x.copy(), polyremoval=pr)) <= TOL
    y = pp.downsample_cF(x, samplerate=500., new_samplerate=100.)
    assert np.array_equal(y, x[::5]) and pp.downsample_cF(x, 500., 100., noCompute=True) == ((201, 7), x.dtype)
    assert pp.detrending_cF(x, polyremoval=None) is x


def test_nan_input_is_refused(engine):
    from syncopy_b200 import _lib
    from syncopy_b200 import preproc as pp
    x = _signal(300, 2, 6)
    x[17, 1] = np.nan
    with pytest.raises(_lib.SpybError):
        pp.sinc_filtering_cF(x, samplerate=500., freq=30.)


def test_batched_filters_match_per_trial(engine):
    import torch
    from syncopy_b200 import preproc as pp
    xs = np.stack([_signal(800, 4, s) for s in range(3)])
    xd = torch.from_numpy(xs).to(engine.tdev)
    y = pp.butterworth_filter(xd, 500., "lp", 40., 4, "twopass", 0, engine).cpu().numpy()
    z = pp.sinc_filter(xd, 500., "lp", 40., 200, "hamming", "onepass", 0, engine).cpu().numpy()
    for k in range(3):
        assert np.array_equal(y[k], pp.but_filtering_cF(xs[k].copy(), 500., "lp", 40., 4, "twopass", 0)[0])
        assert np.array_equal(z[k], pp.sinc_filtering_cF(xs[k].copy(), 500., "lp", 40., 200, "hamming", "onepass", 0)[0])
