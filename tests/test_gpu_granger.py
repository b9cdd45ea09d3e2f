"""GPU parity of the Granger path (K7 regularisation, K8 Wilson factorisation, K9 Granger) through the C ABI."""
import numpy as np
import pytest
import torch

from conftest import load_golden, nerr
from oracle import connectivity as oc
from oracle import synth

pytestmark = pytest.mark.gpu


def mvar_csd(n_chan, n_freq, seed=0, coupling=0.25):
    """Spectral matrix of a stable random MVAR(2) process on n_freq one-sided frequencies (complex128)."""
    rng = np.random.default_rng(seed)
    a1 = 0.45 * np.eye(n_chan) + coupling / np.sqrt(n_chan) * rng.normal(size=(n_chan, n_chan))
    a2 = -0.5 * np.eye(n_chan) + 0.3 * coupling / np.sqrt(n_chan) * rng.normal(size=(n_chan, n_chan))
    q = rng.normal(size=(n_chan, n_chan))
    sigma = q @ q.T / n_chan + np.eye(n_chan)
    om = np.pi * np.arange(n_freq) / (n_freq - 1)
    eye = np.eye(n_chan)
    Hf = np.linalg.inv(eye[None] - a1[None] * np.exp(-1j * om)[:, None, None]
                       - a2[None] * np.exp(-2j * om)[:, None, None])
    S = Hf @ sigma[None] @ Hf.conj().transpose(0, 2, 1)
    return 0.5 * (S + S.conj().transpose(0, 2, 1))


def test_golden_granger_cf(engine):
    """Trial-averaged CSD of the reference -> granger_cF, against the real reference's outputs."""
    from syncopy_b200 import compute_functions as cf
    z, prm = load_golden("granger_ar2_40trials")
    shape, dtype = cf.granger_cF(z["csd_av"][None], noCompute=True)
    assert shape == (1,) + z["csd_av"].shape and dtype == np.float32
    got, meta = cf.granger_cF(z["csd_av"][None], rtol=prm["rtol"], nIter=prm["nIter"], cond_max=prm["cond_max"])
    assert got.shape == shape and got.dtype == np.float32
    assert nerr(got[0], z["granger"].astype(np.float32)) <= 1e-5
    assert bool(meta["converged--bool"]) == bool(z["converged"])
    assert abs(float(meta["max rel. err--float"]) - float(z["err"])) <= 1e-6 * float(z["err"]) + 1e-12
    assert meta["reg. factor--float"] == z["reg_factor"] == 0
    # the reference's condition number comes from a single-precision SVD
    assert abs(float(meta["initial cond. num--float"]) - float(z["cond0"])) <= 1e-4 * float(z["cond0"])
    assert set(meta) == {"converged--bool", "max rel. err--float", "reg. factor--float", "initial cond. num--float"}


def test_golden_wilson_factors(engine):
    z, prm = load_golden("granger_ar2_40trials")
    csd = torch.from_numpy(z["csd_av"]).to(engine.tdev)
    reg, factor, cond0 = engine.regularize_csd(csd, cond_max=prm["cond_max"], eps_max=1e-1)
    assert factor == 0 and torch.equal(reg, csd.to(torch.complex128))
    H, Sigma, conv, err, iters = engine.wilson_sf(reg, n_iter=prm["nIter"], rtol=prm["rtol"])
    assert conv and 1 <= iters <= prm["nIter"]
    assert nerr(H.cpu().numpy(), z["H"]) <= 1e-9
    assert nerr(Sigma.cpu().numpy(), z["Sigma"]) <= 1e-9


def test_golden_regularize_rank1(engine):
    z, prm = load_golden("regularize_rank1")
    csd = torch.from_numpy(z["csd"]).to(engine.tdev)
    reg, factor, cond0 = engine.regularize_csd(csd, cond_max=prm["cond_max"], eps_max=prm["eps_max"])
    assert factor == float(z["eps"])
    assert nerr(reg.cpu().numpy(), z["reg"]) <= 1e-14
    # rank-deficient input: the reference's single-precision SVD saturates near 1/eps32, FP64 goes further
    assert cond0 >= prm["cond_max"] and float(z["cond0"]) >= prm["cond_max"]


@pytest.mark.parametrize("n_chan,cond_max,scale", [(3, 1e3, 1.), (6, 5., 1e-2), (17, 8., 1e-3), (40, 20., 1e-2),
                                                   (12, 3., 1e-4)])
def test_regularize_vs_oracle(engine, n_chan, cond_max, scale):
    """no regularisation needed / eps_max itself / intermediate ladder steps (one within 0.3 % of cond_max)"""
    S = (mvar_csd(n_chan, 20, seed=n_chan) * scale).astype(np.complex64)
    want_reg, want_eps, want_c0 = oc.regularize_csd(S, cond_max=cond_max, eps_max=1e-1)
    reg, factor, cond0 = engine.regularize_csd(torch.from_numpy(S).to(engine.tdev), cond_max=cond_max, eps_max=1e-1)
    assert abs(cond0 - float(want_c0)) <= 2e-4 * float(want_c0)
    assert factor == want_eps
    assert nerr(reg.cpu().numpy(), want_reg.astype(np.complex128)) <= 1e-14


def test_regularize_failure_reports_minus_one(engine):
    a = np.random.default_rng(1).normal(size=(8, 5)) + 0j
    S = (a[:, :, None] * a[:, None, :].conj()).astype(np.complex64) * 1e3        # rank 1, large
    want_reg, want_eps, _ = oc.regularize_csd(S, cond_max=10., eps_max=1e-3)
    reg, factor, _ = engine.regularize_csd(torch.from_numpy(S).to(engine.tdev), cond_max=10., eps_max=1e-3)
    assert want_eps == -1 and factor == -1
    assert nerr(reg.cpu().numpy(), want_reg) <= 1e-14


# (channels, one-sided frequencies): mirrored lengths 32 (radix 16+2), 50 (2*5*5), 46 (2*23, generic prime),
# 128, 500 (4*5^3), 42 (2*3*7), 256
@pytest.mark.parametrize("n_chan,n_freq", [(1, 9), (2, 17), (5, 26), (17, 24), (16, 65), (33, 251), (40, 22),
                                           (64, 129), (130, 33)])
def test_wilson_vs_oracle(engine, n_chan, n_freq):
    S = mvar_csd(n_chan, n_freq, seed=n_chan + n_freq)
    want_H, want_Sig, want_conv, want_err = oc.wilson_sf(S, nIter=60, rtol=1e-9)
    H, Sigma, conv, err, iters = engine.wilson_sf(torch.from_numpy(S).to(engine.tdev), n_iter=60, rtol=1e-9)
    assert conv == want_conv
    assert nerr(H.cpu().numpy(), want_H) <= 1e-8
    assert nerr(Sigma.cpu().numpy(), want_Sig) <= 1e-8
    assert err < 1e-9          # both sit at the rounding floor here; the values themselves are noise
    G = engine.granger(torch.from_numpy(S).to(engine.tdev), H, Sigma).cpu().numpy()
    want_G = oc.granger(S, want_H, want_Sig)
    assert nerr(G, want_G.astype(np.float32)) <= 1e-5


def test_wilson_iteration_cap(engine):
    """nIter exhausted: converged False and the last error are a normal return (wilson_sf.py:109-120)."""
    S = mvar_csd(4, 33, seed=3)
    want_H, want_Sig, want_conv, want_err = oc.wilson_sf(S, nIter=2, rtol=1e-14)
    H, Sigma, conv, err, iters = engine.wilson_sf(torch.from_numpy(S).to(engine.tdev), n_iter=2, rtol=1e-14)
    assert conv is False and want_conv is False and iters == 2
    assert nerr(H.cpu().numpy(), want_H) <= 1e-9
    assert abs(err - want_err) <= 1e-6 * want_err


def test_wilson_not_positive_definite_raises(engine):
    from syncopy_b200._lib import SpybError
    S = mvar_csd(3, 9, seed=1)
    S[4] -= 50 * np.eye(3)
    with pytest.raises(SpybError, match="positive definite"):
        engine.wilson_sf(torch.from_numpy(S).to(engine.tdev))


def test_batched_granger_ar2(engine):
    """
    cfg-4 chain on the reference's AR(2) test network: causality 2 -> 1 only (tests/backend/test_conn.py:245-310).

    With demean_taper=True the DC bin of the CSD is rounding noise (|C_ii(0)| ~ 4e-33 in the reference, float64
    noise squared), and the factorisation depends on the *structure* of that noise matrix: swapping it for another
    noise matrix moves the reference's own low-frequency Granger values by O(1e-2).  Parity is
    therefore checked stage by stage on identical inputs: CSD average vs oracle, then the oracle's granger_cF fed
    with the GPU's CSD average vs the GPU's Granger stage.
    """
    from syncopy_b200 import batched
    trials = synth.ar2_network(40, n_samples=500)
    kw = dict(taper="dpss", taper_opt={"NW": 2.0, "Kmax": 3}, polyremoval=0)
    G, meta, freqs = batched.granger(trials, 200., to_host=True, **kw)
    csd_gpu, _ = batched.cross_spectra(trials, 200., demean_taper=True, to_host=True, **kw)
    av = oc.trial_average([oc.cross_spectra_cF(t.copy(), 200., demean_taper=True, **kw)[0] for t in trials])
    assert nerr(csd_gpu, av) <= 1e-5
    want, want_meta = oc.granger_cF(csd_gpu)
    assert G.shape == want.shape and G.dtype == np.float32
    assert nerr(G, want.astype(np.float32)) <= 1e-5
    assert bool(meta["converged--bool"]) == bool(want_meta["converged--bool"])
    assert meta["reg. factor--float"] == want_meta["reg. factor--float"]
    # against the all-reference chain: same physics away from the noise-dominated lowest bins
    ref_G, _ = oc.granger_cF(av)
    peak = np.argmin(np.abs(freqs - 40.))
    assert G[0, peak, 1, 0] > 0.5 and G[0, peak, 0, 1] < 0.05
    assert abs(G[0, peak, 1, 0] - ref_G[0, peak, 1, 0]) <= 1e-3 * ref_G[0, peak, 1, 0]


def test_full_size_factorisation_property(engine):
    """cfg-4 shape (2049 frequencies x 128 channels): S = H Sigma H^H to the reported error, causal H(0) real."""
    n_chan, n_freq = 128, 2049
    S = torch.from_numpy(mvar_csd(n_chan, n_freq, seed=7)).to(engine.tdev)
    H, Sigma, conv, err, iters = engine.wilson_sf(S, n_iter=100, rtol=5e-6)
    assert conv and err < 5e-6
    Sc = Sigma.to(torch.complex128)
    R = H @ Sc[None] @ H.conj().transpose(1, 2)
    rel = ((S - R).abs() / S.abs()).max().item()
    assert rel <= 2 * 5e-6
    assert torch.equal(Sigma, Sigma.T) or nerr(Sigma.cpu().numpy(), Sigma.T.cpu().numpy()) <= 1e-12
    G = engine.granger(S, H, Sc.real.contiguous())
    assert torch.isfinite(G).all() and G.diagonal(dim1=1, dim2=2).abs().max().item() <= 1e-10


def test_sharded_entry_point_with_trivial_exchange(engine):
    """spyb_wilson_sharded with the whole axis as the slab and a no-op exchange == spyb_wilson (callback plumbing,
    row bookkeeping); the callback sees the lag-domain array (2(nF-1) rows) and the error scalar every iteration."""
    S = torch.from_numpy(mvar_csd(6, 33, seed=5)).to(engine.tdev)
    want = engine.wilson_sf(S, n_iter=40, rtol=1e-9)
    seen = []

    def exchange(what, buf, row_bytes, n_rows):
        seen.append((what, row_bytes, n_rows))
        return 0
    got = engine.wilson_sf(S, n_iter=40, rtol=1e-9, slab=(0, 33), exchange=exchange)
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1]) and got[2:] == want[2:]
    # (1, 8, 1) first and last: the ranks agree on "nobody failed" after the Cholesky and after the loop
    assert seen[0] == (1, 8, 1) and seen[1] == (0, 21 * 16, 64) and seen[2] == (1, 8, 1)
    assert seen[-1] == (1, 8, 1) and len(seen) == 2 * got[4] + 2
