"""GPU parity of the cross-spectral path (K1 -> K2 -> K3) through the C ABI vs oracle / golden."""
import numpy as np
import pytest

from conftest import load_golden, nerr
from oracle import connectivity as oc
from oracle import synth

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("name", ["csd_hann_n512", "csd_dpss_n600_pad1024"])
def test_golden_single_trial_csd(engine, name):
    from syncopy_b200 import compute_functions as cf
    z, prm = load_golden(name)
    kw = prm["kw"]
    got, meta = cf.cross_spectra_cF(z["x"].copy(), prm["fs"], nSamples=kw.get("nSamples"), taper=kw["taper"],
                                    taper_opt=kw.get("taper_opt"), demean_taper=kw.get("demean_taper", False),
                                    polyremoval=None)
    assert got.shape == (1,) + z["csd"].shape and got.dtype == np.complex64
    assert nerr(got[0], z["csd"]) <= TOL
    # Hermitian with exactly real auto-spectra
    assert np.array_equal(got[0], got[0].conj().transpose(0, 2, 1))


@pytest.mark.parametrize("n,c,taper,opt", [(256, 1, "hann", None), (300, 3, "hann", None),
                                           (512, 64, "hann", None), (500, 65, "dpss", {"NW": 3, "Kmax": 5}),
                                           (1024, 130, "hann", None), (128, 200, "dpss", {"NW": 2, "Kmax": 3})])
def test_single_trial_vs_oracle(engine, n, c, taper, opt):
    from syncopy_b200 import compute_functions as cf
    x = np.random.default_rng(c).normal(size=(n, c)).astype("f4")
    foi = np.fft.rfftfreq(n, 1e-3)[3:40]
    got, _ = cf.cross_spectra_cF(x.copy(), 1000., foi=foi, taper=taper, taper_opt=opt, demean_taper=True)
    want, _ = oc.cross_spectra_cF(x.copy(), 1000., foi=foi, taper=taper, taper_opt=opt, demean_taper=True)
    assert got.shape == want.shape
    assert nerr(got, want) <= TOL


def test_golden_coherence_chain(engine):
    """12 trials -> trial-averaged CSD -> coherence, against the real reference's outputs."""
    from syncopy_b200 import batched
    z, prm = load_golden("coherence_12trials")
    csd, freqs = batched.cross_spectra(z["x"], prm["fs"], taper=prm["taper"], polyremoval=None, to_host=True)
    assert csd.shape == (1,) + z["csd_av"].shape
    assert nerr(csd[0], z["csd_av"]) <= TOL
    for output in ("abs", "pow", "fourier", "imag", "real"):
        coh, _ = batched.coherence(z["x"], prm["fs"], taper=prm["taper"], polyremoval=None, output=output,
                                   to_host=True)
        assert coh.dtype == z["coh_" + output].dtype
        assert nerr(coh, z["coh_" + output]) <= TOL


def test_normalize_cf_all_outputs(engine):
    from syncopy_b200 import compute_functions as cf
    z, _ = load_golden("coherence_12trials")
    for output in ("abs", "pow", "fourier", "complex", "angle", "imag", "real", "absreal", "absimag"):
        got = cf.normalize_csd_cF(z["csd_av"][None], output)
        want = oc.normalize_csd_cF(z["csd_av"][None], output)
        assert got.dtype == want.dtype and got.shape == want.shape
        if output == "angle":
            d = np.angle(np.exp(1j * (got.astype(np.float64) - want)))
            assert np.max(np.abs(d) * np.abs(oc.normalize_csd(z["csd_av"][None], "abs"))) <= 1e-5
        else:
            assert nerr(got, want) <= TOL


def test_dyadic_product_cf(engine):
    from syncopy_b200 import compute_functions as cf
    rng = np.random.default_rng(2)
    specs = (rng.normal(size=(4, 3, 20, 7)) + 1j * rng.normal(size=(4, 3, 20, 7))).astype(np.complex64)
    got = cf.spectral_dyadic_product_cF(specs)
    want = oc.spectral_dyadic_product_cF(specs)
    assert got.shape == want.shape and nerr(got, want) <= TOL
    send, rec = [0, 2, 5], [1, 6]
    got = cf.spectral_dyadic_product_cF(specs, send, 3, rec, 2)
    want = oc.spectral_dyadic_product_cF(specs, send, 3, rec, 2)
    assert got.shape == want.shape == (4, 20, 3, 2) and nerr(got, want) <= TOL


def test_batched_equals_per_trial_sum(engine):
    """Trial sharding invariant: CSD sum over all trials == sum of CSD sums over disjoint shards."""
    from syncopy_b200 import batched
    trials = synth.white_noise(9, 256, 12)
    full = batched.cross_spectra_sum(trials, 500., taper="dpss", taper_opt={"NW": 2, "Kmax": 3}, polyremoval=0)
    a = batched.cross_spectra_sum(trials[:4], 500., taper="dpss", taper_opt={"NW": 2, "Kmax": 3}, polyremoval=0)
    b = batched.cross_spectra_sum(trials[4:], 500., taper="dpss", taper_opt={"NW": 2, "Kmax": 3}, polyremoval=0)
    s = (a.csd_sum + b.csd_sum).cpu().numpy()
    assert nerr(full.csd_sum.cpu().numpy(), s) <= 2e-6
    want = oc.trial_average([oc.cross_spectra_cF(t.copy(), 500., taper="dpss", taper_opt={"NW": 2, "Kmax": 3},
                                                 polyremoval=0)[0] for t in trials])
    assert nerr(full.average().cpu().numpy(), want) <= TOL
    kept, _ = batched.cross_spectra(trials[:3], 500., taper="hann", keeptrials=True, to_host=True)
    for k in range(3):
        one = oc.cross_spectra_cF(trials[k].copy(), 500., taper="hann")[0]
        assert nerr(kept[k:k + 1], one) <= TOL


def test_reference_coherence_property(engine):
    """tests/backend/test_conn.py:15-85 on the GPU path: 40 Hz coherence peak between shifted harmonics."""
    from syncopy_b200 import batched
    rng = np.random.default_rng(synth.TEST_SEED)
    n, fs, f = 1001, 1000, 40
    t = np.arange(n) / fs
    shifts = np.array([0, np.pi / 2, np.pi])
    trials = np.stack([np.array([np.cos(f * 2 * np.pi * t + ps) for ps in shifts]).T
                       + rng.standard_normal((n, 3)) for _ in range(100)]).astype("f4")
    coh, freqs = batched.coherence(trials, fs, taper="hann", polyremoval=None, to_host=True)
    c01 = coh[0, :, 0, 1]
    k = np.argmax(c01)
    assert f - 5 < freqs[k] < f + 5 and 0.9 < c01[k] < 1
    assert np.all(c01[:k - 2] < 0.4) and np.all(c01[k + 2:] < 0.4)


def test_cfg2_shape_properties(engine):
    """BASELINE cfg-2 shape (256 ch x 4096 smp), 8 trials: Hermitian, PSD diagonal, shard additivity."""
    from syncopy_b200 import batched
    trials = synth.white_noise(8, 4096, 256)
    res = batched.cross_spectra_sum(trials, 1000., taper="hann", polyremoval=0)
    S = res.csd_sum
    assert S.shape == (2049, 256, 256)
    assert (S - S.conj().transpose(1, 2)).abs().max().item() == 0.0
    d = S.diagonal(dim1=1, dim2=2)
    assert d.imag.abs().max().item() == 0.0 and d.real.min().item() > 0
    # one full frequency against the oracle's definition from GPU spectra-free math
    want = oc.trial_average([oc.cross_spectra_cF(t.copy(), 1000., taper="hann", polyremoval=0,
                                                 foi=np.array([100.0]))[0] for t in trials])
    k = int(np.argmin(np.abs(res.freqs - 100.0)))
    got = (S[k] / 8).cpu().numpy()
    assert nerr(got, want[0, 0]) <= 1e-5


@pytest.mark.parametrize("n_trials,n,c,taper,opt", [(5, 256, 256, "hann", None),
                                                    (3, 128, 128, "dpss", {"NW": 2, "Kmax": 3}),
                                                    (11, 64, 256, "hann", None)])
def test_tcgen05_csd_vs_oracle(engine, n_trials, n, c, taper, opt):
    """Tensor-core contraction (3xTF32, TMEM accumulators) against the oracle's complex64 outer product."""
    import torch
    from syncopy_b200 import batched
    assert engine.csd_planar_supported(c)
    trials = synth.white_noise(n_trials, n, c)
    # wide dynamic range across channels; the loud ones form one FFT channel pair (the mtmfft kernel packs
    # channels (2c, 2c+1) into one complex FFT, whose split error scales with the louder of the two)
    trials[:, :, 2:4] *= 300.0
    trials[:, :, 5] = trials[:, :, 2] * 1e-3 + trials[:, :, 5]
    res = batched.cross_spectra_sum(trials, 1000., taper=taper, taper_opt=opt, polyremoval=0, impl=2)
    simt = batched.cross_spectra_sum(trials, 1000., taper=taper, taper_opt=opt, polyremoval=0, impl=1)
    S = res.csd_sum
    assert (S - S.conj().transpose(1, 2)).abs().max().item() == 0.0
    assert S.diagonal(dim1=1, dim2=2).imag.abs().max().item() == 0.0
    want = oc.trial_average([oc.cross_spectra_cF(t.copy(), 1000., taper=taper, taper_opt=opt, polyremoval=0)[0]
                             for t in trials])
    got = res.average().cpu().numpy()
    # per-frequency normwise error, and error relative to sqrt(S_ii S_jj) (what coherence sees)
    w = want[0].astype(np.complex128)
    err = np.abs(got[0] - w)
    assert (err.max(axis=(1, 2)) / np.abs(w).max(axis=(1, 2))).max() <= TOL
    d = np.sqrt(np.abs(np.einsum("fii->fi", w)))
    assert (err / (d[:, :, None] * d[:, None, :])).max() <= TOL
    assert nerr(S.cpu().numpy(), simt.csd_sum.cpu().numpy()) <= 2e-6
    # accumulate-into (beta = 1) path used for trial chunks
    twice = batched.cross_spectra_sum(trials, 1000., taper=taper, taper_opt=opt, polyremoval=0, impl=2,
                                      out=S.clone())
    assert nerr(twice.csd_sum.cpu().numpy(), 2 * S.cpu().numpy()) <= 1e-6
    assert torch.isfinite(twice.csd_sum.abs()).all()


# ---------------------------------------------------------------------------------------------------------
# "tile slot" path: upper tiles only, per-frequency-slab ownership, sum over source ranks + normalisation
# ---------------------------------------------------------------------------------------------------------
def _planar_spectra(engine, n_trials, n_samples, n_chan, taper, opt, seed=5):
    import torch
    from syncopy_b200 import hostmath as hm
    x = torch.from_numpy(np.random.default_rng(seed).normal(size=(n_trials, n_samples, n_chan)).astype("f4")).to(engine.tdev)
    tapers = engine.taper_table(taper, n_samples, n_samples, opt)
    planes = engine.mtmfft(x, tapers, n_samples, hm.mtmfft_scale(n_samples, n_samples), polyremoval=0,
                           output="fourier_planar", keeptapers=True, freq_major=True)
    return planes, tapers.shape[0]


@pytest.mark.parametrize("n_chan", [128, 256])
@pytest.mark.parametrize("output", ["abs", "pow", "fourier", "imag", "angle"])
def test_tile_path_matches_planar_path(engine, n_chan, output):
    """one rank: accumulate_tiles + normalize_tiles == accumulate_planar + normalize (same sums, mirrored output)"""
    import torch
    planes, K = _planar_spectra(engine, 9, 256, n_chan, "dpss", {"NW": 2, "Kmax": 3})
    nF = planes.shape[0]
    csd = engine.csd_accumulate_planar(planes, alpha=1.0 / K)
    want = engine.csd_normalize(csd[None], output=output, pre_scale=1.0 / 9)[0]
    slots = torch.zeros((1, nF, engine.csd_tile_count(n_chan), 128, 128), dtype=torch.complex64, device=engine.tdev)
    engine.csd_accumulate_tiles(planes, [slots.data_ptr()], [0, nF], 0, alpha=1.0 / K)
    got = engine.csd_normalize_tiles(slots, n_chan, output=output, pre_scale=1.0 / 9)
    assert got.shape == want.shape and got.dtype == want.dtype
    if output == "angle":       # phases of numerically zero imaginary parts on the diagonal: compare as unit vectors
        assert nerr(torch.polar(torch.ones_like(got), got).cpu().numpy(),
                    torch.polar(torch.ones_like(want), want).cpu().numpy()) <= 1e-5
    else:
        assert nerr(got.cpu().numpy(), want.cpu().numpy()) <= 2e-6
    if output in ("abs", "pow"):
        assert torch.equal(got, got.transpose(1, 2))


def test_tile_path_two_sources_and_slabs(engine):
    """two source 'ranks' (trial shards) x two owners (frequency slabs), all on one GPU, chunked accumulation"""
    import torch
    n_chan, T = 256, 10
    planes, K = _planar_spectra(engine, T, 128, n_chan, "hann", None, seed=9)
    nF = planes.shape[0]
    want = engine.csd_normalize(engine.csd_accumulate_planar(planes)[None], output="fourier", pre_scale=1.0 / T)[0]
    f_begin = [0, 30, nF]
    nt = engine.csd_tile_count(n_chan)
    slabs = [torch.zeros((2, f_begin[o + 1] - f_begin[o], nt, 128, 128), dtype=torch.complex64, device=engine.tdev)
             for o in range(2)]
    ptrs = [s.data_ptr() for s in slabs]
    rows = planes.shape[1]
    half = (rows // 2)
    # source 0: first half of the rows in two chunks (beta = 1 accumulates); source 1: the rest
    engine.csd_accumulate_tiles(planes[:, :2], ptrs, f_begin, 0)
    engine.csd_accumulate_tiles(planes[:, 2:half], ptrs, f_begin, 0, beta=1.0)
    engine.csd_accumulate_tiles(planes[:, half:], ptrs, f_begin, 1)
    got = torch.cat([engine.csd_normalize_tiles(slabs[o], n_chan, output="fourier", pre_scale=1.0 / T)
                     for o in range(2)], dim=0)
    assert nerr(got.cpu().numpy(), want.cpu().numpy()) <= 2e-6
    assert (got - got.conj().transpose(1, 2)).abs().max().item() == 0.0


@pytest.mark.parametrize("n_chan", [128, 256])
def test_batched_coherence_tiles_vs_oracle(engine, n_chan):
    from syncopy_b200 import batched
    trials = synth.white_noise(5, 200, n_chan)
    coh, freqs = batched.coherence(trials, 500., taper="hann", polyremoval=0, to_host=True)
    av = oc.trial_average([oc.cross_spectra_cF(t.copy(), 500., taper="hann", polyremoval=0)[0] for t in trials])
    assert coh.shape == (1, 101, n_chan, n_chan)
    assert nerr(coh, oc.normalize_csd(av, "abs")) <= TOL


def test_multi_gpu_tile_exchange_parity():
    """Sharded coherence over 2 GPUs (peer tile stores) == single-GPU result; needs two visible GPUs."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run on a multi-GPU box: gpurun --gpus 2)")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multi_gpu_check.py")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", script],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.parametrize("n_chan", [128, 256])
@pytest.mark.parametrize("output", ["abs", "pow", "fourier", "real", "imag", "angle", "absreal", "absimag"])
def test_fused_coherence_matches_two_kernel_path(engine, n_chan, output):
    """contraction with normalising epilogue == accumulate_planar + normalize, every output conversion"""
    import torch
    planes, K = _planar_spectra(engine, 7, 256, n_chan, "dpss", {"NW": 2, "Kmax": 3}, seed=11)
    want = engine.csd_normalize(engine.csd_accumulate_planar(planes, alpha=1.0 / K)[None], output=output,
                                pre_scale=1.0 / 7)[0]
    got = engine.csd_coherence_planar(planes, output=output)
    assert got.shape == want.shape and got.dtype == want.dtype
    if output == "angle":
        assert nerr(torch.polar(torch.ones_like(got), got).cpu().numpy(),
                    torch.polar(torch.ones_like(want), want).cpu().numpy()) <= 1e-5
    else:
        assert nerr(got.cpu().numpy(), want.cpu().numpy()) <= 3e-6
    if output in ("abs", "pow", "real", "absreal", "absimag"):
        assert torch.equal(got, got.transpose(1, 2))
    if output == "fourier":
        assert (got - got.conj().transpose(1, 2)).abs().max().item() == 0.0
    if output in ("imag", "angle"):
        assert (got + got.transpose(1, 2)).abs().max().item() == 0.0


def test_fused_coherence_many_frequencies(engine):
    """more frequencies than SMs (several per CTA, uneven split) and a row count that is not a multiple of 16"""
    import torch
    torch.manual_seed(3)
    nF, R, C = 333, 37, 256
    planes = torch.randn((nF, R, 2, C), device=engine.tdev)
    want = engine.csd_normalize(engine.csd_accumulate_planar(planes)[None], output="abs", pre_scale=1.0 / R)[0]
    got = engine.csd_coherence_planar(planes, output="abs")
    assert nerr(got.cpu().numpy(), want.cpu().numpy()) <= 3e-6


def test_analog_route_equals_spectral_route(engine):
    """tests/test_connectivity.py:434-473: coherence from AnalogData (cross_spectra_cF) equals the route over
    complex spectra (mtmfft_cF output='fourier' -> spectral_dyadic_product_cF), trial average and normalisation
    included."""
    from syncopy_b200 import compute_functions as cf
    trials = synth.white_noise(6, 400, 10)
    fs = 400.
    foi = np.fft.rfftfreq(400, 1 / fs)
    mk = dict(samplerate=fs, nSamples=None, taper="dpss", taper_opt={"NW": 2, "Kmax": 3})
    a = [cf.cross_spectra_cF(t.copy(), fs, taper="dpss", taper_opt={"NW": 2, "Kmax": 3}, polyremoval=0)[0]
         for t in trials]
    b = []
    for t in trials:
        spec, _ = cf.mtmfft_cF(t.copy(), foi=foi, polyremoval=0, output="fourier", keeptapers=True, method_kwargs=mk)
        b.append(cf.spectral_dyadic_product_cF(spec))
    coh_a = cf.normalize_csd_cF(oc.trial_average(a), "abs")
    coh_b = cf.normalize_csd_cF(oc.trial_average(b), "abs")
    assert coh_a.shape == coh_b.shape == (1, 201, 10, 10)
    assert nerr(coh_a, coh_b) <= 1e-5


def test_csd_mirror_upper(engine):
    """spyb_csd_mirror_upper: lower triangle <- conj(upper), real diagonal, upper triangle untouched"""
    import torch
    torch.manual_seed(3)
    for n_freq, n in ((5, 48), (3, 130)):
        a = torch.randn((n_freq, n, n), dtype=torch.complex64, device=engine.tdev)
        want = torch.triu(a, 1) + torch.triu(a, 1).conj().transpose(1, 2) + torch.diag_embed(torch.diagonal(a, dim1=1, dim2=2).real).to(a.dtype)
        got = engine.csd_mirror_upper(a.clone())
        assert torch.equal(got, want)
        assert torch.equal(got, got.conj().transpose(1, 2))
