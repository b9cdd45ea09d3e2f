"""
GPU parity of the SURVEY 8(f) rows 2-3: jackknife replicates / bias / variance (statistics/jackknifing.py:14-184 behind
connectivity_analysis.py:601-606,736-757), pairwise phase consistency (ST_compRoutines.py:158-233 +
connectivity_analysis.py:624-667) and the single-trial cross-covariance (ST_compRoutines.py:465-584), all against
the oracle restatements that tests/test_oracle_vs_reference.py pins to the reference's own function bodies.

Tolerances: 1e-5 normwise for every directly computed quantity (coherence, replicates, PPC, cross-covariance).  The
jackknife bias multiplies the difference of two O(1) float32 arrays by (T - 1); both sides carry 1e-7 of rounding
there, so the bias is held to (T - 1) * 2e-6 of the coherence scale and the variance to 5e-5 normwise.
"""
import numpy as np
import pytest

from conftest import nerr
from oracle import connectivity as oc
from oracle import statistics as ost
from oracle import synth

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _single_trial_csd(trials, fs, **kw):
    return [oc.cross_spectra_cF(t.copy(), fs, **kw)[0][0] for t in trials]            # [nF, C, C] per trial


@pytest.mark.parametrize("output", ["abs", "pow", "fourier"])
def test_jackknife_coherence(engine, output):
    from syncopy_b200 import statistics as st
    T = 9
    rng = np.random.default_rng(1)
    common = rng.normal(size=(T, 256, 1)).astype("f4")
    trials = (synth.white_noise(T, 256, 5) + 0.8 * common).astype("f4")         # coherent channels
    kw = dict(taper="dpss", taper_opt={"NW": 2, "Kmax": 3}, polyremoval=0)
    direct, bias, var, freqs = st.jackknife_coherence(trials, 500., output=output, to_host=True, **kw)
    d0, b0, v0, reps = ost.jackknife_coherence(_single_trial_csd(trials, 500., **kw), output)
    assert direct.shape == (1,) + d0.shape and direct.dtype == d0.dtype
    assert bias.dtype == b0.dtype and var.dtype == np.float32
    scale = float(np.abs(d0).max())
    assert nerr(direct[0], d0) <= TOL
    assert float(np.abs(bias[0] - b0).max()) <= (T - 1) * 2e-6 * scale
    assert nerr(var[0], v0) <= 5e-5
    # the jackknife estimate direct - bias must agree to working precision as well
    assert nerr(direct[0] - bias[0], d0 - b0) <= (T - 1) * 2e-6


def test_replicates_and_bias_var_generic(engine):
    import torch
    from syncopy_b200 import statistics as st
    rng = np.random.default_rng(3)
    x = (rng.normal(size=(6, 4, 8, 8)) + 1j * rng.normal(size=(6, 4, 8, 8))).astype(np.complex64)
    reps = st.trial_avg_replicates(x)
    want = ost.trial_avg_replicates(list(x))
    assert nerr(reps.cpu().numpy(), want) <= TOL
    direct = ost.trial_mean(list(x))
    b, v = st.bias_var(torch.from_numpy(direct).to(engine.tdev), torch.from_numpy(want).to(engine.tdev))
    b0, v0 = ost.bias_var(direct, list(want))
    assert float(np.abs(b.cpu().numpy() - b0).max()) <= 5 * 2e-6 * float(np.abs(direct).max())
    assert nerr(v.cpu().numpy(), v0) <= 5e-5
    xr = rng.normal(size=(5, 3, 16)).astype("f4")
    assert nerr(st.trial_avg_replicates(xr).cpu().numpy(), ost.trial_avg_replicates(list(xr))) <= TOL


def test_jackknife_granger(engine):
    """
    T + 1 factorisations: direct Granger estimate, bias and variance of the leave-one-out replicates (AR(2) network).
    With demean_taper=True the DC bin of every CSD is rounding noise whose structure steers the factorisation
    (DESIGN.md section 2), so -- like tests/test_gpu_granger.py -- the chain is checked on identical inputs: the GPU's
    single-trial cross spectra go through the GPU jackknife and through the oracle's.
    """
    from syncopy_b200 import batched
    from syncopy_b200 import statistics as st
    T = 8
    trials = synth.ar2_network(T, n_samples=400)
    kw = dict(taper="dpss", taper_opt={"NW": 3, "Kmax": 5}, polyremoval=0)
    cs_gpu, freqs = batched.cross_spectra(trials, 200., demean_taper=True, keeptrials=True, to_host=True, **kw)
    cs_ref = np.stack(_single_trial_csd(trials, 200., demean_taper=True, **kw))
    assert nerr(cs_gpu, cs_ref) <= TOL
    direct, bias, var = st.jackknife_csd(cs_gpu, method="granger", to_host=True)
    d0, b0, v0, reps = ost.jackknife_granger(list(cs_gpu))
    assert direct.shape == (1,) + d0.shape and direct.dtype == np.float32
    assert nerr(direct[0], d0) <= 1e-4
    scale = float(np.abs(d0).max())
    assert float(np.abs(bias[0] - b0).max()) <= (T - 1) * 1e-4 * scale
    assert nerr(var[0], v0) <= 2e-3
    # the chain from the trials themselves: same physics (causality 1 -> 2 only at 40 Hz)
    d_chain, _, v_chain, f2 = st.jackknife_granger(trials, 200., to_host=True, **kw)
    k40 = int(np.argmin(np.abs(f2 - 40.0)))
    assert d_chain[0][k40, 1, 0] > 10 * d_chain[0][k40, 0, 1] and np.isfinite(v_chain).all()
    assert abs(d_chain[0][k40, 1, 0] - d0[k40, 1, 0]) <= 1e-2 * d0[k40, 1, 0]


@pytest.mark.parametrize("T,C", [(2, 3), (7, 6), (20, 16)])
def test_ppc(engine, T, C):
    from syncopy_b200 import statistics as st
    rng = np.random.default_rng(T)
    common = rng.normal(size=(T, 300, 1)).astype("f4")
    trials = (synth.white_noise(T, 300, C) + 0.5 * common).astype("f4")
    kw = dict(taper="hann", polyremoval=0)
    got, freqs = st.ppc(trials, 500., to_host=True, **kw)
    want = ost.ppc(_single_trial_csd(trials, 500., **kw))[None]
    assert got.shape == want.shape == (1, 151, C, C) and got.dtype == np.float32
    assert nerr(got, want) <= TOL
    assert np.abs(np.einsum("fii->fi", got[0]) - 1).max() <= 1e-6       # auto-spectra are in phase with themselves


def test_ppc_column_cf(engine):
    from syncopy_b200 import compute_functions as cf
    rng = np.random.default_rng(6)
    a = (rng.normal(size=(1, 20, 4, 4)) + 1j * rng.normal(size=(1, 20, 4, 4))).astype(np.complex64)
    b = (rng.normal(size=(1, 20, 4, 4)) + 1j * rng.normal(size=(1, 20, 4, 4))).astype(np.complex64)
    got = cf.ppc_column_cF(a, cross_spectrum2=b)
    assert got.shape == a.shape and got.dtype == np.float32
    assert float(np.abs(got - ost.ppc_column_cF(a, b)).max()) <= 2e-6
    assert cf.ppc_column_cF(a, noCompute=True) == (a.shape, np.float32)


@pytest.mark.parametrize("n,c,pr,norm", [(100, 4, 0, False), (101, 3, 1, True), (64, 2, None, False), (1000, 8, 0, True),
                                          (4096, 16, 0, False), (999, 5, 1, False)])
def test_cross_covariance_cf(engine, n, c, pr, norm):
    from syncopy_b200 import compute_functions as cf
    rng = np.random.default_rng(n + c)
    x = rng.normal(size=(n, c)).astype("f4")
    x[:, 1:] += 0.6 * np.roll(x[:, :1], 3, axis=0)                   # lagged coupling: asymmetric cross-covariance
    x += np.float32(0.05)
    got, lags = cf.cross_covariance_cF(x.copy(), samplerate=250., polyremoval=pr, norm=norm, fullOutput=True)
    want, lags0 = ost.cross_covariance_cF(x.copy(), samplerate=250., polyremoval=pr, norm=norm, fullOutput=True)
    assert got.shape == want.shape and got.dtype == np.float32 and np.array_equal(lags, lags0)
    assert nerr(got, want) <= TOL
    assert cf.cross_covariance_cF(x, noCompute=True) == (want.shape, np.float32)
