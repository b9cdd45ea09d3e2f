"""
NumPy model of the in-place decimation-in-frequency FFT of csrc/mtm_dif.cu: pass structure (radices 16, ..., 16,
2^(log2N % 4)), the per-pass twiddle tables of csrc/plan.cu (`make_dif_twiddles`), the digit-reversed output position
`dif_pos`, and the XOR swizzle of the shared-memory layout.  Pure host-side checks of the index arithmetic the kernel
relies on; the kernel itself is tested against the oracle on the GPU.
"""
import numpy as np
import pytest


def radices(log2n):
    return [16] * (log2n // 4) + ([1 << (log2n % 4)] if log2n % 4 else [])


def dif_twiddles(log2n):
    """plan.cu make_dif_twiddles: per radix-16 pass with stride > 1, W_{16*stride}^{o*q} at [(q-1)*stride + o]."""
    tables, stride = [], 1 << log2n
    for _ in range(log2n // 4):
        stride //= 16
        if stride > 1:
            q = np.arange(1, 16)[:, None]
            o = np.arange(stride)[None, :]
            tables.append(np.exp(-2j * np.pi * ((o * q) % (16 * stride)) / (16 * stride)).reshape(-1))
        else:
            tables.append(None)
    return tables


def dif_pos(k, log2n):
    pos, sh = 0, log2n
    for _ in range(log2n // 4):
        sh -= 4
        pos |= (k & 15) << sh
        k >>= 4
    return pos | k


def swz(i):
    return i ^ ((i >> 4) & 15) ^ ((i >> 8) & 15) ^ ((i >> 12) & 15)


def dif_fft(x, log2n):
    n = 1 << log2n
    s = x.astype(np.complex128).copy()
    tables = dif_twiddles(log2n)
    stride = n
    for i, r in enumerate(radices(log2n)):
        stride //= r
        u = np.arange(n // r)
        o = u % stride
        base = (u // stride) * (r * stride) + o
        idx = base[:, None] + np.arange(r)[None, :] * stride          # in place: read and write the same slots
        y = np.fft.fft(s[idx], axis=1)
        if stride > 1:
            q = np.arange(1, r)
            y[:, 1:] *= tables[i][(q[None, :] - 1) * stride + o[:, None]]
        s[idx] = y
    return s


@pytest.mark.parametrize("log2n", [4, 5, 8, 9, 10, 11, 12, 13, 14])
def test_dif_passes_and_output_positions(log2n):
    n = 1 << log2n
    rng = np.random.default_rng(log2n)
    x = rng.normal(size=n) + 1j * rng.normal(size=n)
    s = dif_fft(x, log2n)
    pos = np.array([dif_pos(k, log2n) for k in range(n)])
    assert sorted(pos) == list(range(n))                              # a permutation
    assert np.abs(s[pos] - np.fft.fft(x)).max() <= 1e-9 * n


@pytest.mark.parametrize("log2n", [8, 10, 12, 14])
def test_swizzle_is_xor_linear_and_conflict_free(log2n):
    n = 1 << log2n
    i = np.arange(n)
    assert sorted(swz(i)) == list(range(n))                           # a permutation of the slots
    a, b = np.random.default_rng(1).integers(0, n, size=(2, 1000))
    assert np.array_equal(swz(a ^ b), swz(a) ^ swz(b))                # GF(2)-linear: slot(base + r*stride) = slot(base) ^ const
    # every access pattern of the kernel touches 16 distinct low nibbles per 16 consecutive work items:
    stride = n
    for r in radices(log2n):
        stride //= r
        u = np.arange(n // r)
        base = (u // stride) * (r * stride) + (u % stride)
        for rr in (0, r - 1):
            slots = swz(base + rr * stride)
            groups = slots[: (slots.size // 16) * 16].reshape(-1, 16) & 15
            assert all(len(set(g)) == 16 for g in groups)
    rows = swz(i).reshape(-1, 16) & 15                                # row copies
    assert all(len(set(g)) == 16 for g in rows)
    kf = np.arange(n // 2)                                            # digit-reversed epilogue: bins k ...
    slots = swz(np.array([dif_pos(int(k), log2n) for k in kf])).reshape(-1, 16) & 15
    assert all(len(set(g)) == 16 for g in slots)
    # ... and N - k, where the borrow into the next digit costs one two-way conflict per 16 bins (ncu: 1.3 % excess)
    slots = swz(np.array([dif_pos(int(k), log2n) for k in (n - kf) % n])).reshape(-1, 16) & 15
    assert all(len(set(g)) >= 15 for g in slots)
