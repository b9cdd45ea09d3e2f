"""Host-side product logic (syncopy_b200.hostmath, the cF dry runs) against the oracle.  CPU only."""
import numpy as np
import pytest

from oracle import connectivity as oc
from oracle import spectral as osp
from syncopy_b200 import compute_functions as cf
from syncopy_b200 import hostmath as hm


@pytest.mark.parametrize("taper,opt,n,npad,periodic", [
    ("hann", None, 1000, 1000, False), (None, None, 333, 512, False),
    ("dpss", {"NW": 4, "Kmax": 7}, 4096, 4096, False), ("dpss", {"NW": 2, "Kmax": 3}, 256, 256, True),
    ("kaiser", {"beta": 3}, 500, 600, False), ("hamming", {}, 64, 64, False),
])
def test_taper_tables(taper, opt, n, npad, periodic):
    got = hm.normalized_tapers(taper, n, npad, opt, periodic)
    o = dict(opt or {})
    if periodic and taper == "dpss":
        o["sym"] = False
    want = osp.taper_table(taper, n, npad, o)
    assert got.shape == want.shape and np.array_equal(got, want)


def test_scales():
    assert hm.mtmfft_scale(1000, 1000) == pytest.approx(float(osp.spectrum_scale(1000.0)))
    assert hm.mtmfft_scale(600, 1024) == pytest.approx(float(osp.spectrum_scale(600 * np.sqrt(1024 / 600))))
    assert hm.mtmfft_scale(600, 1024, ft_compat=True) == pytest.approx(float(osp.spectrum_scale(1024)))
    assert hm.stft_scale(256) == pytest.approx(float(osp.spectrum_scale(256)))


def test_best_match_matches_oracle():
    rng = np.random.default_rng(0)
    src = np.fft.rfftfreq(1000, 1e-3)
    for _ in range(50):
        sel = rng.uniform(-10, 600, size=rng.integers(1, 40))
        for squash in (False, True):
            a = hm.best_match(src, sel, squash_duplicates=squash)
            b = osp.best_match(src, sel, squash_duplicates=squash)
            assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])
    # exact ties go right, out-of-range snaps to the ends
    assert list(hm.best_match(np.arange(10.), [2.5, -3, 99])[1]) == [3, 0, 9]
    assert hm.best_match(np.arange(10.), 4.2)[1][0] == 4


def test_polyremoval_codes():
    assert [hm.polyremoval_code(v) for v in (None, False, 0, 1, True, 2)] == [-1, 0, 0, 1, 1, -1]


def test_dry_runs_match_oracle():
    x = np.zeros((1024, 32), dtype=np.float32)
    foi = np.fft.rfftfreq(1024, 1 / 1024.)[10:200]
    for keeptapers, opt in ((True, {}), (True, {"NW": 3, "Kmax": 5}), (False, {"NW": 3, "Kmax": 5})):
        kw = dict(samplerate=1024., nSamples=None, taper="dpss" if opt else "hann", taper_opt=opt)
        for output in ("pow", "fourier", "abs"):
            a = cf.mtmfft_cF(x, foi=foi, keeptapers=keeptapers, output=output, noCompute=True, method_kwargs=kw)
            b = osp.mtmfft_cF(x, foi=foi, keeptapers=keeptapers, output=output, noCompute=True, method_kwargs=kw)
            assert a == b
    # timeAxis=1 (channel-major trial)
    a = cf.cross_spectra_cF(x.T, 1000., nSamples=2048, foi=foi, timeAxis=1, noCompute=True)
    b = oc.cross_spectra_cF(x.T, 1000., nSamples=2048, foi=foi, timeAxis=1, noCompute=True)
    assert a == b
    specs = np.zeros((3, 2, 17, 5), dtype=np.complex64)
    assert cf.spectral_dyadic_product_cF(specs, noCompute=True) == oc.spectral_dyadic_product_cF(specs, noCompute=True)
    assert cf.spectral_dyadic_product_cF(specs, [0, 1], 2, [2, 3, 4], 3, noCompute=True) == \
        oc.spectral_dyadic_product_cF(specs, [0, 1], 2, [2, 3, 4], 3, noCompute=True)
    av = np.zeros((1, 9, 4, 4), dtype=np.complex64)
    for output in ("abs", "fourier", "complex", "angle"):
        assert cf.normalize_csd_cF(av, output, noCompute=True) == oc.normalize_csd_cF(av, output, noCompute=True)
    kw = dict(samplerate=1000., nperseg=128, noverlap=64, taper="hann", taper_opt={})
    for toi in (0.5, np.linspace(0, 1, 7)):
        a = cf.mtmconvol_cF(x, slice(None), slice(None), toi=toi, foi=foi[:20], noCompute=True, method_kwargs=kw)
        b = osp.mtmconvol_cF(x, slice(None), slice(None), toi=toi, foi=foi[:20], noCompute=True, method_kwargs=kw)
        assert a == b
