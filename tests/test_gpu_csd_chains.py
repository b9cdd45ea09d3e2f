"""
GPU parity of the tcgen05 cross-spectral contraction on the row counts the benchmark and production shapes run:
the kernel cuts accumulation chains every 64 (trial x taper) rows, flips its two TMEM buffers and sums the chains
in registers -- rows in {64, 65, 128, 200, 1400} cover 1, 2, 2, 4 and 22 chains.  All three store modes
(`spyb_csd_accumulate_planar`, `spyb_csd_accumulate_tiles`, `spyb_csd_coherence_planar`) are compared with

  * the oracle's arithmetic for the same rows: complex64 outer product + mean over rows
    (syncopy/connectivity/csd.py:98-102, trial sum computational_routine.py:1022-1032), and
  * a float64 contraction (shows the GPU is not the less accurate side).

Tolerance: 1e-5 normwise (north star), per frequency.
"""
import numpy as np
import pytest

from conftest import nerr
from oracle import connectivity as oc
from oracle import spectral as osp
from oracle import synth

pytestmark = pytest.mark.gpu
TOL = 1e-5
ROWS = [64, 65, 128, 200, 1400]


def _planes(n_freq, rows, n_chan, seed):
    """Random planar spectra [nF, R, 2, C] float32 with a channel-gain spread of 1e3 (+ the complex64 view)."""
    rng = np.random.default_rng(seed)
    gains = np.exp(rng.uniform(np.log(0.03), np.log(30.0), size=n_chan)).astype("f4")
    p = rng.normal(size=(n_freq, rows, 2, n_chan)).astype("f4") * gains
    # a coherent component so that off-diagonal entries are not pure noise
    common = rng.normal(size=(n_freq, rows, 2, 1)).astype("f4")
    p[..., : n_chan // 2] += common * gains[: n_chan // 2]
    z = (p[:, :, 0, :] + 1j * p[:, :, 1, :]).astype(np.complex64)
    return p, z


def _oracle_mean(z):
    """csd.py:98-102 on rows: complex64 products, complex64 mean over the row axis -> [nF, C, C]."""
    out = np.empty((z.shape[0], z.shape[2], z.shape[2]), dtype=np.complex64)
    for f in range(z.shape[0]):
        prod = z[f][:, :, None] * z[f][:, None, :].conj()         # complex64
        out[f] = prod.mean(axis=0)
    return out


def _f64_mean(z):
    zz = z.astype(np.complex128)
    return np.einsum("fri,frj->fij", zz, zz.conj()) / z.shape[1]


def _per_freq_err(got, want):
    err = np.abs(got.astype(np.complex128) - want).max(axis=(1, 2))
    return float((err / np.abs(want).max(axis=(1, 2))).max())


@pytest.mark.parametrize("n_chan", [128, 256])
@pytest.mark.parametrize("rows", ROWS)
def test_accumulate_planar_chains(engine, rows, n_chan):
    import torch
    p, z = _planes(5, rows, n_chan, seed=rows + n_chan)
    planes = torch.from_numpy(p).to(engine.tdev)
    got = engine.csd_accumulate_planar(planes, alpha=1.0 / rows).cpu().numpy()
    assert np.array_equal(got, got.conj().transpose(0, 2, 1))
    assert _per_freq_err(got, _oracle_mean(z)) <= TOL
    assert _per_freq_err(got, _f64_mean(z)) <= TOL
    # the same rows in two unequal chunks (beta = 1): chain boundaries fall elsewhere
    cut = rows // 3 + 1
    acc = engine.csd_accumulate_planar(planes[:, :cut].contiguous(), alpha=1.0 / rows)
    acc = engine.csd_accumulate_planar(planes[:, cut:].contiguous(), acc=acc, alpha=1.0 / rows, beta=1.0)
    assert _per_freq_err(acc.cpu().numpy(), _f64_mean(z)) <= TOL


@pytest.mark.parametrize("n_chan", [128, 256])
@pytest.mark.parametrize("rows", ROWS)
def test_accumulate_tiles_chains(engine, rows, n_chan):
    import torch
    p, z = _planes(6, rows, n_chan, seed=3 * rows + n_chan)
    planes = torch.from_numpy(p).to(engine.tdev)
    nF = planes.shape[0]
    nt = engine.csd_tile_count(n_chan)
    # two owners (frequency slabs), one source; rows in two chunks with beta = 1
    f_begin = [0, 2, nF]
    slabs = [torch.zeros((1, f_begin[o + 1] - f_begin[o], nt, 128, 128), dtype=torch.complex64, device=engine.tdev)
             for o in range(2)]
    ptrs = [s.data_ptr() for s in slabs]
    cut = max(1, rows // 2 - 3)
    engine.csd_accumulate_tiles(planes[:, :cut].contiguous(), ptrs, f_begin, 0)
    engine.csd_accumulate_tiles(planes[:, cut:].contiguous(), ptrs, f_begin, 0, beta=1.0)
    want = _f64_mean(z)
    got = torch.cat([engine.csd_normalize_tiles(slabs[o], n_chan, output="fourier", pre_scale=1.0)
                     for o in range(2)], dim=0).cpu().numpy()
    d = np.sqrt(np.abs(np.einsum("fii->fi", want)))
    coh = want / (d[:, :, None] * d[:, None, :])
    assert _per_freq_err(got, coh) <= TOL
    assert np.array_equal(got, got.conj().transpose(0, 2, 1))
    # and against the oracle's own normalisation of its complex64 mean
    assert nerr(got, oc.normalize_csd(_oracle_mean(z), "fourier")) <= TOL


@pytest.mark.parametrize("n_chan", [128, 256])
@pytest.mark.parametrize("rows", ROWS)
@pytest.mark.parametrize("output", ["abs", "fourier"])
def test_coherence_planar_chains(engine, rows, n_chan, output):
    """store mode 3 (the kernel bench.py times at N = 1): contraction + normalising epilogue"""
    import torch
    # more frequencies than one CTA gets in a single round so that TMEM buffers flip between frequencies
    p, z = _planes(5, rows, n_chan, seed=7 * rows + n_chan)
    planes = torch.from_numpy(p).to(engine.tdev)
    got = engine.csd_coherence_planar(planes, output=output).cpu().numpy()
    want64 = _f64_mean(z)
    d = np.sqrt(np.abs(np.einsum("fii->fi", want64)))
    coh64 = want64 / (d[:, :, None] * d[:, None, :])
    ref = oc.normalize_csd(_oracle_mean(z), output)
    if output == "abs":
        coh64 = np.abs(coh64)
    assert got.dtype == ref.dtype and got.shape == ref.shape
    assert _per_freq_err(got, coh64) <= TOL
    assert nerr(got, ref) <= TOL


def _oracle_coherence_subset(trials, fs, taper, taper_opt, fsel):
    """
    The reference chain restricted to the bins `fsel`: per trial mtmfft (complex64 spectra, mtmfft.py:104-127) ->
    complex64 outer product + taper mean (csd.py:98-102) -> sequential complex64 trial sum and division
    (computational_routine.py:1022-1032) -> normalize_csd.  The outer product is per frequency, so restricting the
    bins does not change a single operation on the kept ones.
    """
    acc = None
    for t in trials:
        dat = osp.detrend_trial(np.array(t), 0)
        specs, _ = osp.mtmfft(dat, fs, None, taper, taper_opt, False)       # [K, nF, C] complex64
        s = specs[:, fsel, :]
        cs = (s[:, :, :, None] * s[:, :, None, :].conj()).mean(axis=0)
        if acc is None:
            acc = cs.copy()
        else:
            acc += cs
    acc /= len(trials)
    return oc.normalize_csd(acc[None], "abs")


@pytest.mark.parametrize("taper,opt", [("hann", None), ("dpss", {"NW": 4, "Kmax": 7})])
def test_cfg2_full_size_vs_oracle(engine, taper, opt):
    """BASELINE cfg-2 at full size (200 trials x 256 ch x 4096 smp; hann = 200 rows, DPSS K=7 = 1400 rows)
    through batched.coherence, against the oracle on 24 bins spread over the spectrum (DC and Nyquist included)."""
    from syncopy_b200 import batched
    trials = synth.white_noise(200, 4096, 256)
    fsel = np.unique(np.concatenate([[0, 1, 2, 2047, 2048], np.linspace(3, 2046, 19).astype(int)]))
    coh, freqs = batched.coherence(trials, 1000., taper=taper, taper_opt=opt, polyremoval=0, to_host=True)
    assert coh.shape == (1, 2049, 256, 256) and coh.dtype == np.float32
    want = _oracle_coherence_subset(trials, 1000., taper, opt, fsel)
    got = coh[:, fsel]
    assert nerr(got, want) <= TOL
    # size-independent properties at full size: symmetric, unit diagonal, bounded by one
    assert np.array_equal(coh[0], coh[0].transpose(0, 2, 1))
    dg = np.einsum("fii->fi", coh[0])
    assert np.abs(dg - 1).max() <= 1e-6
    assert coh.max() <= 1 + 1e-6


@pytest.mark.parametrize("n_chan", [64, 96, 192, 320, 384, 512])
@pytest.mark.parametrize("rows", [37, 200])
def test_wide_channel_counts(engine, rows, n_chan):
    """1 to 4 blocks of 128 channels (1 / 3 / 6 / 10 upper tiles), the last block zero-padded when the channel count
    is not a multiple of 128 (64, 96, 192, 320): all three store modes against the float64 contraction"""
    import torch
    nb = (n_chan + 127) // 128
    assert engine.csd_planar_supported(n_chan) and engine.csd_tile_count(n_chan) == nb * (nb + 1) // 2
    assert not engine.csd_planar_supported(48) and not engine.csd_planar_supported(200)
    p, z = _planes(4, rows, n_chan, seed=rows + n_chan)
    planes = torch.from_numpy(p).to(engine.tdev)
    want = _f64_mean(z)
    d = np.sqrt(np.abs(np.einsum("fii->fi", want)))
    coh = want / (d[:, :, None] * d[:, None, :])
    # full Hermitian matrix
    got = engine.csd_accumulate_planar(planes, alpha=1.0 / rows).cpu().numpy()
    assert np.array_equal(got, got.conj().transpose(0, 2, 1))
    assert _per_freq_err(got, want) <= TOL and _per_freq_err(got, _oracle_mean(z)) <= TOL
    # fused coherence
    fused = engine.csd_coherence_planar(planes, output="fourier").cpu().numpy()
    assert _per_freq_err(fused, coh) <= TOL
    assert np.array_equal(engine.csd_coherence_planar(planes, output="abs").cpu().numpy(),
                          engine.csd_coherence_planar(planes, output="abs").cpu().numpy().transpose(0, 2, 1))
    # tile slots + normalisation
    nt = engine.csd_tile_count(n_chan)
    slots = torch.zeros((1, 4, nt, 128, 128), dtype=torch.complex64, device=engine.tdev)
    engine.csd_accumulate_tiles(planes, [slots.data_ptr()], [0, 4], 0)
    tiles = engine.csd_normalize_tiles(slots, n_chan, output="fourier", pre_scale=1.0).cpu().numpy()
    assert _per_freq_err(tiles, coh) <= TOL
    assert np.array_equal(tiles, tiles.conj().transpose(0, 2, 1))


@pytest.mark.parametrize("n_chan", [192, 384])
def test_batched_coherence_wide(engine, n_chan):
    from syncopy_b200 import batched
    trials = synth.white_noise(6, 256, n_chan)
    coh, freqs = batched.coherence(trials, 500., taper="hann", polyremoval=0, to_host=True)
    av = oc.trial_average([oc.cross_spectra_cF(t.copy(), 500., taper="hann", polyremoval=0)[0] for t in trials])
    assert coh.shape == (1, 129, n_chan, n_chan)
    assert nerr(coh, oc.normalize_csd(av, "abs")) <= TOL
    # padded tiles must not leak into the result buffer: a second, differently sized call reuses cached buffers
    assert np.isfinite(coh).all()
