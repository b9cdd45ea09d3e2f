// Packed FP32 arithmetic for the FFT kernels (sm_100a): one 64-bit register pair = one complex number (re, im), and
// add.f32x2 / sub.f32x2 / mul.f32x2 / fma.f32x2 (SASS FADD2 / FMUL2 / FFMA2) work on both halves at once.  ptxas folds
// the re<->im swaps and the sign flips of the helper functions below into operand modifiers, so a complex add is one
// instruction and a complex multiply two.
#pragma once
#include <cuda_runtime.h>

namespace spyb {

typedef unsigned long long c2;

__device__ __forceinline__ c2 pk(float a, float b) {
    c2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float re(c2 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    (void)b;
    return a;
}
__device__ __forceinline__ float im(c2 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    (void)a;
    return b;
}
__device__ __forceinline__ c2 add2(c2 a, c2 b) { c2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ c2 sub2(c2 a, c2 b) { c2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ c2 mul2(c2 a, c2 b) { c2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ c2 fma2(c2 a, c2 b, c2 c) {
    c2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ c2 bc(float s) { return pk(s, s); }
__device__ __forceinline__ c2 mul_mi(c2 x) { return pk(im(x), -re(x)); }                       // x * (-i)
__device__ __forceinline__ c2 cmul2(c2 x, float wx, float wy) {                                // x * (wx + i wy)
    return fma2(pk(-im(x), re(x)), bc(wy), mul2(x, bc(wx)));
}

__device__ __forceinline__ void dft4(c2& a0, c2& a1, c2& a2, c2& a3) {
    const c2 s02 = add2(a0, a2), d02 = sub2(a0, a2), s13 = add2(a1, a3), d13 = sub2(a1, a3);
    a0 = add2(s02, s13);
    a2 = sub2(s02, s13);
    const c2 t = mul_mi(d13);          // X1 = d02 - i d13, X3 = d02 + i d13
    a1 = add2(d02, t);
    a3 = sub2(d02, t);
}

// forward 16-point DFT; X[k] ends up in x[reg16(k)] (same conventions as Radix<16> in fft_core.cuh)
__host__ __device__ constexpr int reg16(int k) { return (k >> 2) + 4 * (k & 3); }
__device__ __forceinline__ void dft16(c2 (&x)[16]) {
    dft4(x[0], x[4], x[8], x[12]);
    dft4(x[1], x[5], x[9], x[13]);
    dft4(x[2], x[6], x[10], x[14]);
    dft4(x[3], x[7], x[11], x[15]);
    const float h = 0.70710678118654752440f;    // cos(pi/4)
    const float c1 = 0.92387953251128675613f;   // cos(pi/8)
    const float s1 = 0.38268343236508977173f;   // sin(pi/8)
    // x[i + 4q] *= W16^{i q}
    x[5] = cmul2(x[5], c1, -s1);                                           // W^1
    x[9] = mul2(add2(x[9], mul_mi(x[9])), bc(h));                          // W^2 = (1 - i)/sqrt2
    x[13] = cmul2(x[13], s1, -c1);                                         // W^3
    x[6] = mul2(add2(x[6], mul_mi(x[6])), bc(h));                          // W^2
    x[10] = mul_mi(x[10]);                                                 // W^4 = -i
    x[14] = mul2(sub2(mul_mi(x[14]), x[14]), bc(h));                       // W^6 = (-1 - i)/sqrt2
    x[7] = cmul2(x[7], s1, -c1);                                           // W^3
    x[11] = mul2(sub2(mul_mi(x[11]), x[11]), bc(h));                       // W^6
    x[15] = cmul2(x[15], -c1, s1);                                         // W^9
    dft4(x[0], x[1], x[2], x[3]);
    dft4(x[4], x[5], x[6], x[7]);
    dft4(x[8], x[9], x[10], x[11]);
    dft4(x[12], x[13], x[14], x[15]);
}

__device__ __forceinline__ c2 cmulc2(c2 x, c2 w) { return cmul2(x, re(w), im(w)); }

// forward 8-point DFT, natural order in and out
__device__ __forceinline__ void dft8(c2 (&x)[8]) {
    // n = i + 2m, k = q + 4s: four-point DFTs over m, twiddles W8^(i q), two-point DFTs over i
    dft4(x[0], x[2], x[4], x[6]);
    dft4(x[1], x[3], x[5], x[7]);
    const float h = 0.70710678118654752440f;
    x[3] = mul2(add2(x[3], mul_mi(x[3])), bc(h));                          // W8^1 = (1 - i)/sqrt2
    x[5] = mul_mi(x[5]);                                                   // W8^2 = -i
    x[7] = mul2(sub2(mul_mi(x[7]), x[7]), bc(h));                          // W8^3 = (-1 - i)/sqrt2
    // X[q] = T0[q] + T1[q], X[q + 4] = T0[q] - T1[q] with T_i[q] in x[i + 2q]
    c2 y[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) { y[q] = add2(x[2 * q], x[2 * q + 1]); y[q + 4] = sub2(x[2 * q], x[2 * q + 1]); }
#pragma unroll
    for (int q = 0; q < 8; ++q) x[q] = y[q];
}

}  // namespace spyb
