"""
TEST INFRASTRUCTURE -- regenerate tests/golden/*.npz from the REAL reference.

Run in the build container (needs /root/reference):

    python -m oracle.make_golden

For every case the unmodified reference backend (imported by file path, see
oracle/ref_loader.py) is run on a seeded input; its output is stored next to
the input, and the oracle restatement is asserted to reproduce it (bit-exact
where the arithmetic is identical, <= 2e-6 normwise otherwise).  The committed
.npz files are what pins the oracle on machines without the reference (the GPU
box).  Reference: esi-neuroscience/syncopy v2023.09 @ a86199a, run under
NumPy/SciPy of this image (recorded in each file).
"""
import json
import os
import sys

import numpy as np
import scipy

from . import connectivity as oc
from . import ref_loader
from . import spectral as osp
from . import synth
from . import timefreq as otf

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def nerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def save(name, params, **arrays):
    meta = dict(params=params, numpy=np.__version__, scipy=scipy.__version__,
                reference="esi-neuroscience/syncopy v2023.09 a86199a (by-path import)")
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, meta=np.array(json.dumps(meta)), **arrays)
    print(f"  wrote {name}.npz ({os.path.getsize(path) / 1024:.0f} KiB)")


def check(tag, got, want, tol):
    e = nerr(got, want)
    status = "ok" if e <= tol else "FAIL"
    print(f"  [{status}] {tag}: oracle vs reference normwise err {e:.2e} (tol {tol:.0e})")
    if e > tol:
        raise SystemExit(f"oracle disagrees with reference on {tag}")


def main():
    ref = ref_loader.load()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    rng = np.random.default_rng(synth.TEST_SEED)

    # ---- mtmfft ------------------------------------------------------------
    print("mtmfft")
    cases = {
        "mtmfft_hann_n1000": dict(n=1000, c=3, fs=1000., kw=dict(taper="hann")),
        "mtmfft_boxcar_odd_n1001": dict(n=1001, c=2, fs=1000., kw=dict(taper=None)),
        "mtmfft_dpss_pad2048": dict(n=1500, c=4, fs=500., kw=dict(
            nSamples=2048, taper="dpss", taper_opt={"NW": 3.0, "Kmax": 5}, demean_taper=True)),
        "mtmfft_kaiser_ftcompat_n512": dict(n=512, c=2, fs=256., kw=dict(
            nSamples=600, taper="kaiser", taper_opt={"beta": 3}, ft_compat=True)),
    }
    for name, cs in cases.items():
        x = (rng.normal(size=(cs["n"], cs["c"])) + 0.3).astype("f4")
        want, fr = ref.mtmfft.mtmfft(x.copy(), cs["fs"], **cs["kw"])
        got, fr2 = osp.mtmfft(x.copy(), cs["fs"], **cs["kw"])
        check(name, got, want, 1e-12)
        assert np.array_equal(fr, fr2)
        save(name, dict(fs=cs["fs"], kw=cs["kw"]), x=x, ftr=want, freqs=fr)

    # ---- csd + normalize_csd -----------------------------------------------
    print("csd")
    cases = {
        "csd_hann_n512": dict(n=512, c=6, fs=1000., kw=dict(taper="hann")),
        "csd_dpss_n600_pad1024": dict(n=600, c=5, fs=200., kw=dict(
            nSamples=1024, taper="dpss", taper_opt={"NW": 2.5, "Kmax": 4}, demean_taper=True)),
    }
    for name, cs in cases.items():
        x = rng.normal(size=(cs["n"], cs["c"])).astype("f4")
        x[:, 1] += 0.5 * x[:, 0]
        want, fr = ref.csd.csd(x.copy(), cs["fs"], **cs["kw"])
        got, _ = oc.csd(x.copy(), cs["fs"], **cs["kw"])
        check(name, got, want, 1e-12)
        save(name, dict(fs=cs["fs"], kw=cs["kw"]), x=x, csd=want, freqs=fr)

    print("normalize_csd / trial average")
    trials = synth.white_noise(12, 400, 4)
    trials[:, :, 2] += 0.7 * trials[:, :, 0]
    per_trial = [ref.csd.csd(t.copy(), 1000., taper="hann")[0] for t in trials]
    av = np.zeros_like(per_trial[0])
    for p in per_trial:                       # computational_routine.py:1025,1032
        av += p
    av /= len(per_trial)
    check("trial_average", oc.trial_average([oc.csd(t.copy(), 1000., taper="hann")[0] for t in trials]), av, 0.0)
    outs = {}
    for output in ("abs", "pow", "fourier", "angle", "imag", "real"):
        want = ref.csd.normalize_csd(av[None], output)
        check(f"normalize_csd[{output}]", oc.normalize_csd(av[None], output), want, 0.0)
        outs["coh_" + output] = want
    save("coherence_12trials", dict(fs=1000., taper="hann"), x=trials, csd_av=av, **outs)

    # ---- stft / mtmconvol --------------------------------------------------
    print("mtmconvol")
    cases = {
        "mtmconvol_dpss_zeros": dict(n=2000, c=4, fs=1000., kw=dict(
            nperseg=256, noverlap=192, taper="dpss", taper_opt={"NW": 2.0, "Kmax": 3},
            boundary="zeros", padded=True, detrend="constant")),
        "mtmconvol_hann_nobdry_linear": dict(n=500, c=2, fs=600., kw=dict(
            nperseg=100, noverlap=99, taper="hann", taper_opt={},
            boundary=None, padded=False, detrend="linear")),
        "mtmconvol_boxcar_nodetrend": dict(n=1000, c=2, fs=500., kw=dict(
            nperseg=128, noverlap=64, taper=None, taper_opt=None,
            boundary="zeros", padded=True, detrend=False)),
    }
    for name, cs in cases.items():
        x = (rng.normal(size=(cs["n"], cs["c"])) + np.linspace(0, 2, cs["n"])[:, None]).astype("f4")
        kw_ref = {k: (dict(v) if isinstance(v, dict) else v) for k, v in cs["kw"].items()}
        want, fr = ref.mtmconvol.mtmconvol(x.copy(), cs["fs"], **kw_ref)
        got, fr2 = osp.mtmconvol(x.copy(), cs["fs"], **cs["kw"])
        check(name, got, want, 1e-12)
        save(name, dict(fs=cs["fs"], kw=cs["kw"]), x=x, ftr=want, freqs=fr)

    # ---- wavelets ----------------------------------------------------------
    print("cwt")
    x = rng.normal(size=(1000, 3)).astype("f4")
    fs = 500.
    for wname, wref, wora in (("morlet6", ref.wavelets_mod.Morlet(6), otf.Morlet(6)),
                              ("paul4", ref.wavelets_mod.Paul(4), otf.Paul(4)),
                              ("dog2", ref.wavelets_mod.DOG(2), otf.DOG(2))):
        foi = np.array([4., 9., 17., 33., 60., 120., 200.])
        scales = wref.scale_from_period(1 / foi)
        assert np.allclose(scales, wora.scale_from_period(1 / foi), rtol=1e-15)
        want = ref.wavelet.wavelet(x.copy(), fs, scales, wref)
        got = otf.wavelet(x.copy(), fs, scales, wora)
        check("cwt_" + wname, got, want, 1e-7)
        save("cwt_" + wname, dict(fs=fs, wavelet=wname), x=x, scales=scales, spec=want)

    # ---- superlets ---------------------------------------------------------
    print("superlet")
    x = rng.normal(size=(600, 2)).astype("f4")
    x[:, 0] += 2 * np.cos(2 * np.pi * 40 * np.arange(600) / 500.)
    fs = 500.
    foi = np.linspace(10, 100, 10)
    scales = ref.superlet.scale_from_period(1 / foi)
    for adaptive in (False, True):
        kw = dict(order_max=5, order_min=1, c_1=3, adaptive=adaptive)
        want = ref.superlet.superlet(x.copy(), fs, scales, **kw)
        got = otf.superlet(x.copy(), fs, scales, **kw)
        name = "superlet_faslt" if adaptive else "superlet_mult"
        check(name, got, want, 1e-6)
        save(name, dict(fs=fs, kw=kw), x=x, scales=scales, spec=want)

    # ---- Wilson / Granger --------------------------------------------------
    print("wilson / granger")
    fs = 200.
    trials = synth.ar2_network(40, n_samples=500)
    per_trial = [ref.csd.csd(t.copy(), fs, taper="dpss", taper_opt={"NW": 2.0, "Kmax": 3},
                             demean_taper=True)[0] for t in trials]
    av = np.zeros_like(per_trial[0])
    for p in per_trial:
        av += p
    av /= len(per_trial)
    reg, eps, cond0 = ref.wilson_sf.regularize_csd(av, cond_max=1e4, eps_max=1e-1)
    reg2, eps2, cond02 = oc.regularize_csd(av, cond_max=1e4, eps_max=1e-1)
    assert eps == eps2 and cond0 == cond02 and np.array_equal(reg, reg2)
    reg = reg.astype(np.complex128)
    H, Sigma, conv, err = ref.wilson_sf.wilson_sf(reg, nIter=100, rtol=5e-6)
    H2, Sigma2, conv2, err2 = oc.wilson_sf(reg, nIter=100, rtol=5e-6)
    check("wilson H", H2, H, 1e-12)
    check("wilson Sigma", Sigma2, Sigma, 1e-12)
    assert conv == conv2 and abs(err - err2) <= 1e-12 * max(err, 1e-300) + 1e-15
    G = ref.granger.granger(reg, H, Sigma)
    check("granger", oc.granger(reg, H2, Sigma2), G, 1e-12)
    save("granger_ar2_40trials", dict(fs=fs, rtol=5e-6, nIter=100, cond_max=1e4),
         x=trials, csd_av=av, H=H, Sigma=Sigma, converged=np.array(conv), err=np.array(err),
         reg_factor=np.array(eps), cond0=np.array(cond0), granger=G)

    # an ill-conditioned CSD that needs the regularisation ladder
    a = rng.normal(size=(30, 6)) + 1j * rng.normal(size=(30, 6))
    bad = (a[:, :, None] * a[:, None, :].conj()).astype(np.complex64)    # rank-1 per frequency
    bad += 1e-7 * np.eye(6, dtype=np.complex64)
    reg, eps, cond0 = ref.wilson_sf.regularize_csd(bad, cond_max=1e4, eps_max=1e-1)
    reg2, eps2, cond02 = oc.regularize_csd(bad, cond_max=1e4, eps_max=1e-1)
    assert eps == eps2 and np.array_equal(reg, reg2)
    save("regularize_rank1", dict(cond_max=1e4, eps_max=1e-1), csd=bad, reg=reg,
         eps=np.array(eps), cond0=np.array(cond0))

    # ---- best_match --------------------------------------------------------
    print("best_match")
    src = np.fft.rfftfreq(1000, 1 / 1000.)
    sel = np.array([0.2, 10.5, 10.4, 499.7, 600., 33.3, 10.6, 1.5])
    for squash in (False, True):
        v, i = ref.tools.best_match(src, sel, squash_duplicates=squash)
        v2, i2 = osp.best_match(src, sel, squash_duplicates=squash)
        assert np.array_equal(i, i2) and np.array_equal(v, v2)
    save("best_match", {}, source=src, selection=sel,
         idx=ref.tools.best_match(src, sel)[1],
         idx_squashed=ref.tools.best_match(src, sel, squash_duplicates=True)[1])
    print("all golden vectors written; oracle == reference on every case")


if __name__ == "__main__":
    sys.exit(main())
