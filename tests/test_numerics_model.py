"""
NumPy model of the operand split the tcgen05 cross-spectral kernel uses (csrc/csd_tc.cu):
    x = hi + lo,  hi = rna_tf32(x),  lo = x - hi;   x_i x_j ~= hi_i hi_j  [TF32 MMA]  +  bf16(hi_i) bf16(lo_j) + bf16(lo_i) bf16(hi_j)  [BF16 MMAs]
It bounds the error the split itself introduces (accumulation idealised in float64), independent of any GPU, and
shows why the cross terms may run at bf16 precision while plain TF32 (or an all-bf16 two-way split) would miss the
1e-5 parity bar.
"""
import numpy as np


def rna_tf32(x):
    """cvt.rna.tf32.f32: keep 10 explicit mantissa bits, round to nearest, ties away from zero."""
    b = np.asarray(x, dtype=np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def bf16(x):
    """cvt.rn.bf16.f32: keep 7 explicit mantissa bits, round to nearest even."""
    b = np.asarray(x, dtype=np.float32).view(np.uint32)
    r = b + np.uint32(0x7FFF) + ((b >> np.uint32(16)) & np.uint32(1))
    return (r & np.uint32(0xFFFF0000)).view(np.float32)


def contraction_errors(spread):
    rng = np.random.default_rng(0)
    R, C = 200, 64
    gain = 10.0 ** np.linspace(-spread / 2, spread / 2, C)
    x = (rng.normal(size=(R, C)) * gain).astype(np.float32)          # one real plane is enough: products are real
    exact = x.astype(np.float64).T @ x.astype(np.float64)
    d = np.sqrt(np.diag(exact))
    scale = d[:, None] * d[None, :]                                   # the scale of the coherence normalisation
    hi = rna_tf32(x)
    lo = (x - hi).astype(np.float32)
    h64, l64 = hi.astype(np.float64), lo.astype(np.float64)
    hb, lb = bf16(hi).astype(np.float64), bf16(lo).astype(np.float64)
    schemes = {
        "tf32 only": h64.T @ h64,
        "3xTF32": h64.T @ h64 + h64.T @ rna_tf32(lo).astype(np.float64) + rna_tf32(lo).astype(np.float64).T @ h64,
        "tf32 + bf16 cross terms": h64.T @ h64 + hb.T @ lb + lb.T @ hb,
        "bf16 two-way split": (lambda a, b: a.T @ a + a.T @ b + b.T @ a)(
            bf16(x).astype(np.float64), bf16(x - bf16(x)).astype(np.float64)),
    }
    return {k: float((np.abs(v - exact) / scale).max()) for k, v in schemes.items()}


def test_split_error_budget():
    for spread in (0, 6):
        e = contraction_errors(spread)
        assert e["tf32 + bf16 cross terms"] <= 6e-7, e           # the shipped scheme: far inside the 1e-5 bar
        assert e["3xTF32"] <= 2e-7, e                             # (on the GPU both sit near 1e-6: FP32 accumulation)
        assert e["tf32 only"] >= 2e-5, e                          # plain TF32 misses the bar
        assert e["bf16 two-way split"] >= 5 * e["tf32 + bf16 cross terms"], e


def test_rounding_helpers():
    x = np.float32([1.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -10, -3.1415927, 1e-30, 6.5e4])
    assert np.all(np.abs(rna_tf32(x) - x) <= np.abs(x) * 2.0 ** -11)
    assert rna_tf32(np.float32(1.0 + 2.0 ** -11)) == np.float32(1.0 + 2.0 ** -10)      # tie rounds away
    assert np.all(np.abs(bf16(x) - x) <= np.abs(x) * 2.0 ** -8)
    assert bf16(np.float32(1.0 + 2.0 ** -8)) == np.float32(1.0)                       # tie rounds to even
