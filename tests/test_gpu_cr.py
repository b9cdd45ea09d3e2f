"""
GPU tests of the batched `compute_sequential` stand-in (syncopy_b200/cr.py) against the reference
runtime's contract (syncopy/shared/computational_routine.py:944-1035): every selection entry k -- permuted, repeated,
of its own length -- must land at `targetLayout[k]` with the values the per-trial cF (oracle) produces, bit-exact in
its placement; `keeptrials=False` = sum in selection order / nTrials.  Reference test for the trial order:
syncopy/tests/test_specest.py:155-158 (`trials=[3, 1, 0]`).
"""
import numpy as np
import pytest

from conftest import nerr
from oracle import connectivity as oc
from oracle import spectral as osp
from oracle import timefreq as otf
from oracle import synth

pytestmark = pytest.mark.gpu
TOL = 1e-5
FS = 500.0


def _dataset(lengths, n_chan, seed=3):
    """AnalogData-like 2-D dataset: trials stacked along time + trialdefinition (start, stop, offset)."""
    rng = np.random.default_rng(seed)
    bounds, off = [], 0
    for n in lengths:
        bounds.append((off, off + n, 0))
        off += n
    data = rng.normal(size=(off, n_chan)).astype("f4")
    for t, (a, b, _) in enumerate(bounds):
        data[a:b] += np.float32(0.01 * t)                # make trials distinguishable
    return data, np.array(bounds)


@pytest.mark.parametrize("trial_ids", [None, [3, 1, 0], [2, 2, 5, 0, 2]])
def test_mtmfft_selection_order_and_unequal_lengths(engine, trial_ids):
    from syncopy_b200 import cr
    lengths = [400, 512, 300, 512, 450, 400]
    data, td = _dataset(lengths, 6)
    nS = 512                                             # pad='maxperlen': all trials padded to the longest
    foi = np.fft.rfftfreq(nS, 1 / FS)
    res = cr.compute_sequential(data, td, "mtmfft", FS, trial_ids=trial_ids, keeptrials=True, nSamples=nS,
                                taper="dpss", taper_opt={"NW": 2, "Kmax": 3}, polyremoval=0, output="pow",
                                keeptapers=True, chunk_bytes=3 * 512 * 6 * 4)          # several chunks per length
    ids = list(range(len(lengths))) if trial_ids is None else trial_ids
    out = res["result"]
    assert out.shape == (len(ids), 3, nS // 2 + 1, 6) and out.dtype == np.float32
    assert [(s.start, s.stop) for s in res["layout"]] == [(k, k + 1) for k in range(len(ids))]
    mk = dict(samplerate=FS, nSamples=nS, taper="dpss", taper_opt={"NW": 2, "Kmax": 3})
    for k, t in enumerate(ids):
        a, b, _ = td[t]
        want, _ = osp.mtmfft_cF(data[a:b].copy(), foi=foi, keeptapers=True, polyremoval=0, output="pow", method_kwargs=mk)
        assert nerr(out[res["layout"][k]], want) <= TOL
    # repeated selection entries are computed independently and must agree bit for bit
    if trial_ids is not None and trial_ids.count(2) > 1:
        pos = [k for k, t in enumerate(ids) if t == 2]
        for p in pos[1:]:
            assert np.array_equal(out[pos[0]], out[p])


def test_mtmfft_trial_average_and_preallocated_target(engine):
    from syncopy_b200 import cr
    data, td = _dataset([256] * 7, 8)
    foi = np.fft.rfftfreq(256, 1 / FS)
    ids = [6, 0, 3, 3]
    target = np.zeros((1, 1, 129, 8), dtype=np.float32)
    res = cr.compute_sequential(data, td, "mtmfft", FS, trial_ids=ids, keeptrials=False, target=target, taper="hann",
                                polyremoval=0, output="pow", keeptapers=False, chunk_bytes=2 * 256 * 8 * 4)
    assert res["result"] is target
    mk = dict(samplerate=FS, nSamples=None, taper="hann", taper_opt={})
    want = oc.trial_average([osp.mtmfft_cF(data[td[t, 0]:td[t, 1]].copy(), foi=foi, keeptapers=False, polyremoval=0,
                                           output="pow", method_kwargs=mk)[0] for t in ids])
    assert nerr(target, want) <= TOL


def test_mtmconvol_rows_of_unequal_trials(engine):
    """time-frequency results of unequal trials occupy different numbers of rows (targetLayout stacking)"""
    from syncopy_b200 import cr
    lengths = [600, 1000, 600, 840]
    data, td = _dataset(lengths, 4)
    ids = [1, 3, 0, 2, 1]
    res = cr.compute_sequential(data, td, "mtmconvol", FS, trial_ids=ids, keeptrials=True, nperseg=100, noverlap=50,
                                taper="hann", polyremoval=0, output="pow", keeptapers=False,
                                chunk_bytes=1000 * 4 * 4)
    out, lay = res["result"], res["layout"]
    rows = [int(np.ceil(lengths[t] / 50)) for t in ids]
    assert [s.stop - s.start for s in lay] == rows and out.shape[0] == sum(rows)
    mk = dict(samplerate=FS, nperseg=100, noverlap=50, taper="hann", taper_opt={})
    foi = np.fft.rfftfreq(100, 1 / FS)
    for k, t in enumerate(ids):
        a, b, _ = td[t]
        want = osp.mtmconvol_cF(data[a:b].copy(), slice(None), slice(None), equidistant=True, toi=0.5, foi=foi,
                                keeptapers=False, polyremoval=0, output="pow", method_kwargs=dict(mk))
        assert out[lay[k]].shape == want.shape
        assert nerr(out[lay[k]], want) <= TOL


def test_wavelet_stacking_from_memmap(engine, tmp_path):
    from syncopy_b200 import cr
    from syncopy_b200 import hostmath as hm
    lengths = [300, 300, 420]
    data, td = _dataset(lengths, 3)
    path = tmp_path / "analog.dat"
    mm = np.memmap(path, dtype="f4", mode="w+", shape=data.shape)
    mm[:] = data
    mm.flush()
    src = np.memmap(path, dtype="f4", mode="r", shape=data.shape)
    wav = hm.Morlet(6)
    foi = np.array([20., 40., 80.])
    scales = wav.scale_from_period(1 / foi)
    ids = [2, 0]
    res = cr.compute_sequential(src, td, "wavelet", FS, trial_ids=ids, keeptrials=True, scales=scales, wavelet=wav,
                                polyremoval=0, output="pow")
    out, lay = res["result"], res["layout"]
    assert out.shape == (420 + 300, 1, 3, 3)
    owav = otf.Morlet(6)
    for k, t in enumerate(ids):
        a, b, _ = td[t]
        want = otf.wavelet_cF(data[a:b].copy(), slice(None), slice(None), toi="all", polyremoval=0, output="pow",
                              method_kwargs=dict(samplerate=FS, scales=owav.scale_from_period(1 / foi), wavelet=owav))
        assert nerr(out[lay[k]], want) <= TOL


@pytest.mark.parametrize("n_chan", [12, 128])
def test_coherence_chain(engine, n_chan):
    from syncopy_b200 import cr
    data, td = _dataset([256] * 9, n_chan)
    ids = [8, 1, 1, 4, 0, 7]
    res = cr.compute_sequential(data, td, "coh", FS, trial_ids=ids, keeptrials=False, taper="dpss",
                                taper_opt={"NW": 2, "Kmax": 3}, polyremoval=0, output="abs",
                                chunk_bytes=2 * 256 * n_chan * 4)
    av = oc.trial_average([oc.cross_spectra_cF(data[td[t, 0]:td[t, 1]].copy(), FS, taper="dpss",
                                               taper_opt={"NW": 2, "Kmax": 3}, polyremoval=0)[0] for t in ids])
    want = oc.normalize_csd(av, "abs")
    assert res["result"].shape == want.shape
    assert nerr(res["result"], want) <= TOL
    assert res["h2d_bytes"] == len(ids) * 256 * n_chan * 4


def test_single_trial_csd_stack(engine):
    from syncopy_b200 import cr
    data, td = _dataset([200, 200, 200], 5)
    ids = [2, 0]
    res = cr.compute_sequential(data, td, "csd", FS, trial_ids=ids, keeptrials=True, taper="hann", polyremoval=0)
    out = res["result"]
    assert out.shape == (2, 101, 5, 5) and out.dtype == np.complex64
    for k, t in enumerate(ids):
        want, _ = oc.cross_spectra_cF(data[td[t, 0]:td[t, 1]].copy(), FS, taper="hann", polyremoval=0)
        assert nerr(out[k:k + 1], want) <= TOL
