"""
TEST INFRASTRUCTURE (see oracle/__init__.py) -- CPU restatement of the
cross-spectral / Granger part of the reference hot path:

    syncopy/connectivity/csd.py:16-115                   -> csd
    syncopy/connectivity/csd.py:118-172                  -> normalize_csd
    syncopy/connectivity/ST_compRoutines.py:268-424      -> cross_spectra_cF
    syncopy/connectivity/ST_compRoutines.py:29-117       -> spectral_dyadic_product_cF
    syncopy/connectivity/AV_compRoutines.py:35-112       -> normalize_csd_cF
    syncopy/connectivity/wilson_sf.py:16-254             -> wilson_sf, regularize_csd
    syncopy/connectivity/granger.py:10-79                -> granger
    syncopy/connectivity/AV_compRoutines.py:292-412      -> granger_cF
    syncopy/shared/computational_routine.py:1017-1032    -> trial_average
"""
import numpy as np

from .spectral import (OUTPUT_DTYPES, best_match, convert_output, detrend_trial,
                       freqs_hash, mtmfft)


def csd(trl_dat, samplerate=1, nSamples=None, taper="hann", taper_opt=None,
        demean_taper=False, norm=False):
    """
    csd.py:16-115.  CS[f, i, j] = mean_k X[k,f,i] * conj(X[k,f,j]), computed
    and averaged in complex64 exactly like the reference (it materialises the
    [K, nFreq, C, C] product and calls .mean(axis=0)).
    """
    specs, freqs = mtmfft(trl_dat, samplerate, nSamples, taper, taper_opt, demean_taper)
    # [K, F, i, j]: first channel index un-conjugated (csd.py:98,102,115 composed)
    prod = specs[:, :, :, None] * specs[:, :, None, :].conj()
    cs = prod.mean(axis=0)
    if norm:
        if taper != "dpss":
            raise ValueError("single-trial normalisation needs taper='dpss' (csd.py:104-108)")
        diag = np.einsum("fii->fi", cs)
        cs = cs / np.sqrt(diag[:, :, None] * diag[:, None, :])
    return cs, freqs


def normalize_csd(csd_av, output="abs"):
    """csd.py:118-172: coherency C_ij / sqrt(C_ii C_jj) (complex sqrt), then conversion."""
    diag = csd_av.diagonal(axis1=-2, axis2=-1)
    denom = np.sqrt(diag[..., None] * diag[..., None, :])
    return convert_output(csd_av / denom, output)


def cross_spectra_cF(trl_dat, samplerate=1, nSamples=None, foi=None, taper="hann",
                     taper_opt=None, demean_taper=False, polyremoval=False, timeAxis=0,
                     chunkShape=None, noCompute=False):
    """ST_compRoutines.py:268-424."""
    dat = trl_dat.T if timeAxis != 0 else trl_dat
    if nSamples is None:
        nSamples = dat.shape[0]
    freqs = np.fft.rfftfreq(nSamples, 1 / samplerate)
    if foi is not None:
        _, fidx = best_match(freqs, foi, squash_duplicates=True)
        n_freq = fidx.size
    else:
        fidx, n_freq = slice(None), freqs.size
    out_shape = (1, n_freq, dat.shape[1], dat.shape[1])
    if noCompute:
        return out_shape, OUTPUT_DTYPES["fourier"]
    dat = detrend_trial(np.array(dat), polyremoval)
    cs, freqs = csd(dat, samplerate, nSamples, taper=taper, taper_opt=taper_opt,
                    demean_taper=demean_taper)
    return cs[None, fidx, ...], {"freqs_hash": freqs_hash(freqs)}


def spectral_dyadic_product_cF(specs, send_idx=None, send_N=None, rec_idx=None, rec_N=None,
                               chunkShape=None, noCompute=False):
    """ST_compRoutines.py:29-117.  specs [nTime, K, nFreq, C] complex64."""
    n_time, _, n_freq, n_chan = specs.shape
    if send_idx is not None:
        shape = (n_time, n_freq, send_N, rec_N)
    else:
        shape = (n_time, n_freq, n_chan, n_chan)
    if noCompute:
        return shape, OUTPUT_DTYPES["fourier"]
    if send_idx is not None:
        prod = specs[..., send_idx, None] * specs[..., None, rec_idx].conj()
    else:
        prod = specs[..., None] * specs[..., None, :].conj()
    return prod.mean(axis=1)


def normalize_csd_cF(csd_av_dat, output="abs", chunkShape=None, noCompute=False):
    """AV_compRoutines.py:35-112."""
    fmt = OUTPUT_DTYPES["fourier"] if output in ("complex", "fourier") else OUTPUT_DTYPES["abs"]
    if noCompute:
        return csd_av_dat.shape, fmt
    return normalize_csd(csd_av_dat, output)


def trial_average(per_trial_results):
    """
    What the runtime does for keeptrials=False
    (computational_routine.py:1022-1032): sequential `+=` into a buffer of the
    result dtype, then `/= nTrials`.
    """
    it = iter(per_trial_results)
    acc = np.array(next(it), copy=True)
    n = 1
    for res in it:
        acc += res
        n += 1
    acc /= n
    return acc


# ---------------------------------------------------------------------------
# Wilson spectral factorisation + Granger
# ---------------------------------------------------------------------------

def max_rel_err(A, B):
    """wilson_sf.py:190-194."""
    return (np.abs(A - B) / np.abs(A)).max()


def regularize_csd(CSD, cond_max=1e3, eps_max=1e-3, nSteps=15):
    """wilson_sf.py:197-254: brute-force `CSD + eps*I` ladder on the max 2-norm condition number."""
    eye = np.eye(CSD.shape[1])
    cond0 = np.linalg.cond(CSD).max()
    if cond0 < cond_max:
        return CSD, 0, cond0
    for eps in np.logspace(-10, np.log10(eps_max), nSteps):
        reg = CSD + eps * eye
        if np.linalg.cond(reg).max() < cond_max:
            return reg, eps, cond0
    return reg, -1, cond0


def _mirror(a, n_freq):
    """attach the negative frequencies: [a, conj(a[nFreq-2:0:-1])] (wilson_sf.py:63,70)."""
    return np.concatenate((a, a[n_freq - 2:0:-1].conj()), axis=0)


def _psi0_initial(CSD):
    """wilson_sf.py:123-151: transpose of chol(Re(sym(gamma_0))), gamma = fft(CSD, axis 0)."""
    gamma0 = np.fft.fft(CSD, axis=0)[0]
    gamma0 = np.real((gamma0 + gamma0.T.conj()) / 2)
    ev = np.linalg.eigvals(gamma0)
    if np.all(np.imag(ev) == 0):
        psi0 = np.linalg.cholesky(gamma0)
    else:
        psi0 = np.ones(gamma0.shape)
    return psi0.T


def _plus_operator(g):
    """wilson_sf.py:154-184: causal projection through the lag domain."""
    n_lag = g.shape[0] // 2
    beta = np.real(np.fft.ifft(g, axis=0))
    beta[0] *= 0.5
    g0 = beta[0].copy()
    beta[n_lag] *= 0.5
    beta[n_lag + 1:] = 0
    return np.fft.fft(beta, axis=0), g0


def wilson_sf(CSD, nIter=100, rtol=1e-6):
    """
    wilson_sf.py:16-120 (direct_inversion=True branch, the only one the cF uses).
    Returns H[:nFreq], Sigma, converged, err.
    """
    n_freq = CSD.shape[0]
    eye = np.eye(CSD.shape[1])
    S = _mirror(CSD, n_freq)
    psi0 = _psi0_initial(S)
    psi = _mirror(np.tile(psi0, (n_freq, 1, 1)), n_freq)
    U = np.linalg.cholesky(S)
    converged, err = False, np.inf
    for _ in range(nIter):
        g = np.linalg.inv(psi) @ U
        g = g @ g.conj().transpose(0, 2, 1)
        gplus, gplus0 = _plus_operator(g + eye)
        A = np.triu(gplus0)
        A = A - A.conj().T
        psi = psi @ (gplus + A)
        psi0 = psi0 @ (gplus0 + A)
        err = max_rel_err(S, psi @ psi.conj().transpose(0, 2, 1))
        if err < rtol:
            converged = True
            break
    Sigma = psi0 @ psi0.T                      # plain transpose (wilson_sf.py:114)
    H = psi @ np.linalg.inv(psi0)
    return H[:n_freq], Sigma, converged, err


def granger(CSD, H, Sigma):
    """
    granger.py:10-79 written index-wise:
        G[f,i,j] = ln( S_jj / (S_jj - (Sig_ii - Sig_ji^2 / Sig_jj) |H_ji|^2) )
    with S_jj = |CSD[f,j,j]|, Sig_ab = |Sigma[a,b]|.
    """
    S = np.abs(np.einsum("fjj->fj", CSD))           # [F, j]
    sig = np.abs(Sigma)
    sig_d = np.abs(np.diag(Sigma))                  # [C]
    Hji2 = np.abs(H.transpose(0, 2, 1)) ** 2        # [F, i, j] = |H[f,j,i]|^2
    # SigmaII[i, j] = sig_d[j]; SigmaII.T[i, j] = sig_d[i]; SigmaJI[i, j] = sig[j, i]
    fac = sig_d[:, None] - sig.T ** 2 / sig_d[None, :]
    Smat = S[:, None, :] * np.ones(CSD.shape[1])[:, None]
    return np.log(Smat / (Smat - fac * Hji2))


def granger_cF(csd_av_dat, rtol=5e-6, nIter=100, cond_max=1e4, chunkShape=None, noCompute=False):
    """AV_compRoutines.py:292-412."""
    if noCompute:
        return csd_av_dat.shape, OUTPUT_DTYPES["abs"]
    CSD = csd_av_dat[0]
    CSDreg, factor, ini_cn = regularize_csd(CSD, cond_max=cond_max, eps_max=1e-1)
    CSDreg = CSDreg.astype(np.complex128)
    H, Sigma, conv, err = wilson_sf(CSDreg, nIter=nIter, rtol=rtol)
    G = granger(CSDreg, H, Sigma)
    meta = {
        "converged--bool": np.array(conv),
        "max rel. err--float": np.array(err),
        "reg. factor--float": np.array(factor),
        "initial cond. num--float": np.array(ini_cn),
    }
    return G[None, ...], meta
