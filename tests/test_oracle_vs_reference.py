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
