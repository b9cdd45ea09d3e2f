"""
NumPy model of the plus operator as csrc/wilson.cu evaluates it (plus_pack / plus_causal / plus_unpack kernels):
only the one-sided frequencies and the upper-triangle matrix elements are transformed, `g_ji = conj g_ij` and
`g(-f) = conj g(f)` supply the rest, and one forward FFT of `w_l (beta_ij[l] + i beta_ij[N - l])` yields both
`gplus_ij` and `gplus_ji`.  Checked against the oracle's restatement of wilson_sf.py:154-184 (full mirrored arrays).
"""
import numpy as np
import pytest

from oracle import connectivity as oc


def plus_operator_one_sided(g):
    """g [nF, C, C] Hermitian per frequency (one-sided) -> (M [nF, C, C] = gplus + S, M0 = gplus_0 + S)."""
    nf, c, _ = g.shape
    length = 2 * (nf - 1)
    iu, ju = np.triu_indices(c)
    # pack: conj(g + I) on the full circle, mirror = conjugate (so the mirrored half holds g + I itself)
    ge = g[:, iu, ju] + (iu == ju)[None, :]
    w = np.empty((length, iu.size), dtype=np.complex128)
    w[:nf] = ge.conj()
    w[nf:] = ge[nf - 2:0:-1]
    # forward FFT of the conjugate = N * conj(ifft); only the real part is used
    beta = np.fft.fft(w, axis=0).real / length
    # causal window, packing beta_ij[l] and beta_ji[l] = beta_ij[N - l]
    wl = np.ones(nf)
    wl[0] = wl[nf - 1] = 0.5
    z = np.zeros((length, iu.size), dtype=np.complex128)
    rev = (length - np.arange(nf)) % length
    z[:nf] = wl[:, None] * (beta[:nf] + 1j * beta[rev])
    g0 = 0.5 * beta[0]
    zf = np.fft.fft(z, axis=0)
    z1 = zf[:nf]
    z2 = zf[(length - np.arange(nf)) % length].conj()
    gij = 0.5 * (z1 + z2)
    gji = (z1 - z2) / 2j
    m = np.zeros((nf, c, c), dtype=np.complex128)
    off = iu != ju
    m[:, iu, ju] = gij + np.where(off, g0, 0.0)[None, :]
    m[:, ju[off], iu[off]] = gji[:, off] - g0[off][None, :]
    m0 = np.zeros((c, c))
    m0[iu, ju] = np.where(off, 2 * g0, g0)
    return m, m0


@pytest.mark.parametrize("n_freq,n_chan", [(5, 1), (9, 3), (26, 4), (33, 6)])
def test_one_sided_plus_operator_matches_reference_formulation(n_freq, n_chan):
    rng = np.random.default_rng(n_freq)
    x = rng.normal(size=(n_freq, n_chan, 2 * n_chan)) + 1j * rng.normal(size=(n_freq, n_chan, 2 * n_chan))
    x[0] = x[0].real
    x[-1] = x[-1].real                                   # DC and Nyquist of a real process are real
    g = x @ x.conj().transpose(0, 2, 1)
    full = oc._mirror(g, n_freq) + np.eye(n_chan)
    gplus, g0 = oc._plus_operator(full)
    a = np.triu(g0)
    a = a - a.conj().T
    want_m = (gplus + a)[:n_freq]
    want_m0 = g0 + a
    m, m0 = plus_operator_one_sided(g)
    assert np.abs(m - want_m).max() <= 1e-12 * np.abs(want_m).max()
    assert np.abs(m0 - want_m0).max() <= 1e-12 * np.abs(want_m0).max()
    # the update matrix of psi0 is upper triangular: psi0 stays triangular through the iteration
    assert np.abs(np.tril(m0, -1)).max() == 0.0
