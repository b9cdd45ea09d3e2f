"""
Host side of the wavelet / superlet kernels, checked on the CPU: the convolution plan the engine uploads
(`hostmath.conv_same_length`, `conv_same_spectrum`, `cwt_taps`, `superlet_taps`, `compute_functions.superlet_tables`)
is evaluated here with NumPy FFTs -- the arithmetic `cwt_kernel` performs on the device -- and must reproduce the
oracle's `fftconvolve(..., 'same')` transforms, including kernels longer than the trial.
"""
import numpy as np
import pytest

from oracle import synth
from oracle import timefreq as otf
from syncopy_b200 import hostmath as hm
from syncopy_b200.compute_functions import superlet_tables


def plan_transform(x, taps_per_scale, exponents):
    """NumPy emulation of Engine.conv_plan + cwt_kernel for one trial x [N, C] -> [nScales, N, C] complex128."""
    n = x.shape[0]
    flat = [t for tl in taps_per_scale for t in tl]
    L = hm.conv_same_length(n, flat)
    X = np.fft.fft(np.concatenate([x.astype(np.float64), np.zeros((L - n, x.shape[1]))]), axis=0)
    out = []
    for tl, el in zip(taps_per_scale, exponents):
        z = np.ones((n, x.shape[1]), dtype=np.complex128)
        for taps, a in zip(tl, el):
            kern = hm.conv_same_spectrum(taps, n, L)             # FFT_L(h) / L
            y = np.fft.ifft(X * (kern * L)[:, None], axis=0)[:n]
            z = z * (y if a == 1.0 else np.abs(y) ** a * np.exp(1j * a * np.angle(y)))
        out.append(z)
    return np.stack(out)


def nerr(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.mark.parametrize("wav_o,wav_g", [(otf.Morlet(6), hm.Morlet(6)), (otf.Paul(4), hm.Paul(4)),
                                         (otf.DOG(2), hm.DOG(2))])
@pytest.mark.parametrize("n", [300, 512])
def test_cwt_plan_reproduces_fftconvolve_same(wav_o, wav_g, n):
    fs = 200.
    x = synth.white_noise_trial(n, 3, 11)
    foi = np.array([1.5, 7., 30., 90.])                 # the lowest one has a support far beyond the trial
    scales = wav_o.scale_from_period(1 / foi)
    if isinstance(wav_g, hm.Morlet):
        assert len(hm.cwt_taps(wav_g, scales[0], 1 / fs)) > n
    want = otf.cwt(x.astype(np.float64), wav_o, scales, 1 / fs)
    got = plan_transform(x, [[hm.cwt_taps(wav_g, s, 1 / fs)] for s in scales], [[1.0]] * scales.size)
    assert got.shape == want.shape
    assert nerr(got, want) <= 1e-6                       # the oracle rounds to complex64


@pytest.mark.parametrize("adaptive", [False, True])
def test_superlet_plan_reproduces_reference_bookkeeping(adaptive):
    fs, n = 500., 400
    x = synth.white_noise_trial(n, 2, 5)
    foi = np.arange(10., 100., 20.)
    scales = 1.0 / (2 * np.pi * foi)
    if adaptive:
        scales = np.sort(scales)[::-1]                   # FASLT wants high -> low scales (freqanalysis.py:940-950)
    kw = dict(order_max=6, order_min=1, c_1=3, adaptive=adaptive)
    want = otf.superlet(x.copy(), fs, scales, **kw)
    taps, expo = superlet_tables(scales, 1 / fs, **kw)
    got = plan_transform(x, taps, expo)
    assert got.shape == want.shape
    # magnitudes: the reference multiplies complex64 powers, so 1e-5; phases only where the magnitude is not tiny
    assert nerr(np.abs(got), np.abs(want)) <= 2e-5
    big = np.abs(want) > 1e-3 * np.abs(want).max()
    assert np.abs(np.angle(got[big] * np.conj(want[big]))).max() <= 1e-3


def test_conv_same_length_is_minimal_power_of_two():
    n = 1000
    for m in (1, 2, 7, 999, 1000, 1001, 5000, 40001):
        taps = np.ones(m)
        L = hm.conv_same_length(n, [taps])
        c0 = (m - 1) // 2
        need = n + max(min(c0, n - 1), min(m - 1 - c0, n - 1))
        assert L >= need and (L & (L - 1)) == 0 and (L // 2 < need or L == 16)


def test_conv_same_length_beyond_the_shared_memory_limit():
    """power of two up to 16384; the smallest 5-smooth multiple of 16 beyond (global-memory FFT path)"""
    import numpy as np
    from syncopy_b200 import hostmath as hm
    assert hm.conv_same_length(8192, [np.zeros(98)]) == 16384
    for n, m in [(20000, 4001), (18000, 301), (100000, 10001), (16385, 3)]:
        L = hm.conv_same_length(n, [np.zeros(m)])
        c0 = (m - 1) // 2
        assert L >= n + max(c0, m - 1 - c0) and L % 16 == 0
        r = L
        for f in (2, 3, 5):
            while r % f == 0:
                r //= f
        assert r == 1
        p2 = 16
        while p2 < L:
            p2 *= 2
        assert L <= p2
    assert hm.conv_same_length(20000, [np.zeros(4001)]) == 23040


def test_conv_groups_overlap_save_reproduces_fftconvolve_same():
    """short kernels as overlap-save blocks (hostmath.conv_groups + conv_segment_spectrum), long ones through the
    full padded length: every scale equals scipy.signal.fftconvolve(x, taps, 'same') (transform.py:88-108)"""
    import scipy.signal as ss
    rng = np.random.default_rng(0)
    n = 3000
    x = rng.normal(size=n)
    taps = [[rng.normal(size=m) + 1j * rng.normal(size=m)] for m in (31, 96, 300, 955, 2100, 4000, 7000)]
    taps[1].append(rng.normal(size=40) + 0j)                       # a scale with two factors of different length
    L = hm.conv_same_length(n, [t for tl in taps for t in tl])
    groups = hm.conv_groups(n, taps, L)
    assert groups[0]["s0"] == 0 and groups[-1]["s1"] == len(taps)
    assert all(a["s1"] == b["s0"] for a, b in zip(groups, groups[1:]))
    assert any(g["seg"] for g in groups) and any(not g["seg"] for g in groups)
    for g in groups:
        for si in range(g["s0"], g["s1"]):
            for t in taps[si]:
                want = ss.fftconvolve(x, t, "same")
                if g["seg"]:
                    Ls, V, A = g["L"], g["V"], g["A"]
                    assert V > 0 and g["n_seg"] * V >= n
                    K = hm.conv_segment_spectrum(t, Ls, A) * Ls
                    got = np.zeros(n, complex)
                    for b in range(g["n_seg"]):
                        lo = b * V - A
                        seg = np.zeros(Ls)
                        a0, a1 = max(lo, 0), min(lo + Ls, n)
                        seg[a0 - lo:a1 - lo] = x[a0:a1]
                        y = np.fft.ifft(np.fft.fft(seg) * K)
                        n1 = min(n, b * V + V)
                        got[b * V:n1] = y[:n1 - b * V]
                else:
                    xp = np.zeros(g["L"])
                    xp[:n] = x
                    got = np.fft.ifft(np.fft.fft(xp) * hm.conv_same_spectrum(t, n, g["L"]) * g["L"])[:n]
                assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


def test_conv_groups_edges():
    """empty time selections and transforms shorter than a block keep the single full-length group; a long trial with
    one short and one long kernel splits into a block run and a full-length run"""
    g = hm.conv_groups(0, [[np.ones(5)]], 16)
    assert len(g) == 1 and not g[0]["seg"] and g[0]["s1"] == 1
    g = hm.conv_groups(10, [[np.ones(5)]], 16)
    assert len(g) == 1 and not g[0]["seg"]
    taps = [[np.ones(31)], [np.ones(20001)]]
    L = hm.conv_same_length(100000, [t for tl in taps for t in tl])
    g = hm.conv_groups(100000, taps, L)
    assert [x["seg"] for x in g] == [True, False] and g[0]["L"] in (4096, 8192) and g[0]["n_seg"] * g[0]["V"] >= 100000
    assert g[1]["L"] == L
