"""
The two host-checkable facts behind the regularisation kernels (csrc/wilson.cu, K7): (1) for a Hermitian matrix the
2-norm condition number np.linalg.cond computes from an SVD equals max|lambda| / min|lambda| of its eigenvalues, and
adding eps*I only shifts the eigenvalues -- so one tridiagonalisation serves the whole eps ladder; (2) the ladder the
library walks reproduces np.logspace(-10, log10(eps_max), nSteps) bit for bit (wilson_sf.py:236-254).
"""
import math

import numpy as np
import pytest


def ladder(eps_max, n_steps):
    """csrc/wilson.cu regularize_csd: y = s * step + start, the last point is `stop` itself, eps = 10**y."""
    stop = math.log10(eps_max)
    step = (stop + 10.0) / (n_steps - 1) if n_steps > 1 else 0.0
    return [10.0 ** (stop if (n_steps > 1 and s == n_steps - 1) else s * step - 10.0) for s in range(n_steps)]


@pytest.mark.parametrize("eps_max,n_steps", [(1e-1, 15), (1e-3, 15), (0.5, 7), (1e-3, 2)])
def test_ladder_matches_numpy_logspace(eps_max, n_steps):
    want = np.logspace(-10, np.log10(eps_max), n_steps)
    got = np.array(ladder(eps_max, n_steps))
    assert np.allclose(got, want, rtol=4e-16, atol=0.0)
    assert got[-1] == want[-1]


@pytest.mark.parametrize("n,rank_deficient", [(3, False), (8, False), (6, True)])
def test_condition_number_from_eigenvalue_extremes(n, rank_deficient):
    rng = np.random.default_rng(n)
    k = 2 if rank_deficient else 2 * n
    a = rng.normal(size=(5, n, k)) + 1j * rng.normal(size=(5, n, k))
    csd = a @ a.conj().transpose(0, 2, 1)
    if rank_deficient:
        csd = csd + 1e-9 * np.eye(n)                     # slightly lifted, as the golden vector is
    for eps in (0.0, 1e-6, 1e-2):
        reg = csd + eps * np.eye(n)
        lam = np.linalg.eigvalsh(csd) + eps              # shifted eigenvalues of the unregularised matrix
        want = np.linalg.cond(reg)
        got = np.abs(lam).max(axis=1) / np.abs(lam).min(axis=1)
        assert np.allclose(got, want, rtol=1e-3 if rank_deficient else 1e-10)   # lambda_min ~ 1e-9 lambda_max: eps64 * cond
