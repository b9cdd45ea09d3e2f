// Register-resident Stockham FFT for one thread block (sm_100a, FP32).
//
// A complex FFT of length N = 2^LOG2N is shared by NT = N/16 threads; every thread keeps
// E = 16 complex values in registers.  At the start of each pass thread j owns the elements
// with indices j + NT*e (e = 0..15).  A pass of radix R (16, or 2/4/8 for the remainder
// pass that comes last) does E/R in-register DFTs, then scatters the results into shared
// memory at their Stockham auto-sort positions; after a barrier every thread re-gathers
// its "j + NT*e" set.  P independent FFTs ("channel pairs", see mtmfft.cu) are interleaved
// lane-wise: shared word address = pad(idx) * P + p, with pad(idx) = idx + idx/16, which
// makes every scatter/gather of the exchange conflict-free for 8-byte words (checked with
// tools/fft_model.py for all N in 2^4..2^14 and P in {1,2,4,8}).
//
// Twiddles come from a per-length table computed on the host in double precision
// (fft_plan.cu): for the pass with stride NS the table holds W_{NS*R}^{k*r} at
// [(r-1)*NS + k], k < NS, r = 1..R-1, so that lanes with consecutive k read consecutive
// words.  Passes are concatenated in execution order (the first pass has NS = 1 and needs
// no twiddles).
#pragma once
#include "common.cuh"
#include "packed.cuh"

namespace spyb {

__host__ __device__ __forceinline__ constexpr int fft_pad(int idx) { return idx + (idx >> 4); }
__host__ __device__ constexpr int fft_padded_len(int n) { return n + (n >> 4); }

// ---------------------------------------------------------------------------------------
// in-register DFTs (forward, e^{-2 pi i nk/R}); result X[k] ends up in register reg_of(k)
// ---------------------------------------------------------------------------------------
template <typename C>
__device__ __forceinline__ void dft2(C& a, C& b) {
    C t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

template <typename C>
__device__ __forceinline__ void dft4(C& a0, C& a1, C& a2, C& a3) {
    C s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
    a0 = cadd(s02, s13);
    a2 = csub(s02, s13);
    // X1 = d02 - i*d13, X3 = d02 + i*d13
    a1.x = d02.x + d13.y; a1.y = d02.y - d13.x;
    a3.x = d02.x - d13.y; a3.y = d02.y + d13.x;
}

template <int R> struct Radix;

template <> struct Radix<2> {
    __host__ __device__ static constexpr int reg_of(int k) { return k; }
    __device__ __forceinline__ static void run(float2 (&x)[2]) { dft2(x[0], x[1]); }
};

template <> struct Radix<4> {
    __host__ __device__ static constexpr int reg_of(int k) { return k; }
    __device__ __forceinline__ static void run(float2 (&x)[4]) { dft4(x[0], x[1], x[2], x[3]); }
};

template <> struct Radix<8> {
    // n = i + 2m, k = q + 4s: T_i[q] in x[i + 2q]; X[q + 4s] in x[2q + s]
    __host__ __device__ static constexpr int reg_of(int k) { return 2 * (k & 3) + (k >> 2); }
    __device__ __forceinline__ static void run(float2 (&x)[8]) {
        dft4(x[0], x[2], x[4], x[6]);
        dft4(x[1], x[3], x[5], x[7]);
        const float h = 0.70710678118654752440f;
        // x[1+2q] *= W8^q
        { float2 t = x[3]; x[3] = make_float2((t.x + t.y) * h, (t.y - t.x) * h); }     // (1-i)/sqrt2
        x[5] = cmul_mi(x[5]);                                                          // -i
        { float2 t = x[7]; x[7] = make_float2((t.y - t.x) * h, -(t.x + t.y) * h); }    // (-1-i)/sqrt2
        dft2(x[0], x[1]);
        dft2(x[2], x[3]);
        dft2(x[4], x[5]);
        dft2(x[6], x[7]);
    }
};

template <> struct Radix<16> {
    // n = i + 4m, k = q + 4s: T_i[q] in x[i + 4q]; X[q + 4s] in x[s + 4q]
    __host__ __device__ static constexpr int reg_of(int k) { return (k >> 2) + 4 * (k & 3); }
    __device__ __forceinline__ static void run(float2 (&x)[16]) {
        dft4(x[0], x[4], x[8], x[12]);
        dft4(x[1], x[5], x[9], x[13]);
        dft4(x[2], x[6], x[10], x[14]);
        dft4(x[3], x[7], x[11], x[15]);
        const float h = 0.70710678118654752440f;    // cos(pi/4)
        const float c1 = 0.92387953251128675613f;   // cos(pi/8)
        const float s1 = 0.38268343236508977173f;   // sin(pi/8)
        // x[i + 4q] *= W16^{i q}
        x[5] = cmul(x[5], make_float2(c1, -s1));                                         // W^1
        { float2 t = x[9]; x[9] = make_float2((t.x + t.y) * h, (t.y - t.x) * h); }       // W^2
        x[13] = cmul(x[13], make_float2(s1, -c1));                                       // W^3
        { float2 t = x[6]; x[6] = make_float2((t.x + t.y) * h, (t.y - t.x) * h); }       // W^2
        x[10] = cmul_mi(x[10]);                                                          // W^4
        { float2 t = x[14]; x[14] = make_float2((t.y - t.x) * h, -(t.x + t.y) * h); }    // W^6
        x[7] = cmul(x[7], make_float2(s1, -c1));                                         // W^3
        { float2 t = x[11]; x[11] = make_float2((t.y - t.x) * h, -(t.x + t.y) * h); }    // W^6
        x[15] = cmul(x[15], make_float2(-c1, s1));                                       // W^9
        dft4(x[0], x[1], x[2], x[3]);
        dft4(x[4], x[5], x[6], x[7]);
        dft4(x[8], x[9], x[10], x[11]);
        dft4(x[12], x[13], x[14], x[15]);
    }
};

// Packed-FP32 versions (packed.cuh: one instruction per complex add, two per complex multiply) with the same
// register conventions as Radix<R>::run above; used by the pass below.
template <int R> struct RadixP;
template <> struct RadixP<2> {
    __device__ __forceinline__ static void run(c2 (&x)[2]) { const c2 t = x[0]; x[0] = add2(t, x[1]); x[1] = sub2(t, x[1]); }
};
template <> struct RadixP<4> {
    __device__ __forceinline__ static void run(c2 (&x)[4]) { dft4(x[0], x[1], x[2], x[3]); }
};
template <> struct RadixP<8> {
    __device__ __forceinline__ static void run(c2 (&x)[8]) {
        dft4(x[0], x[2], x[4], x[6]);
        dft4(x[1], x[3], x[5], x[7]);
        const float h = 0.70710678118654752440f;
        x[3] = mul2(add2(x[3], mul_mi(x[3])), bc(h));          // (1 - i)/sqrt2
        x[5] = mul_mi(x[5]);                                   // -i
        x[7] = mul2(sub2(mul_mi(x[7]), x[7]), bc(h));          // (-1 - i)/sqrt2
#pragma unroll
        for (int q = 0; q < 4; ++q) { const c2 t = x[2 * q]; x[2 * q] = add2(t, x[2 * q + 1]); x[2 * q + 1] = sub2(t, x[2 * q + 1]); }
    }
};
template <> struct RadixP<16> {
    __device__ __forceinline__ static void run(c2 (&x)[16]) { dft16(x); }
};

// ---------------------------------------------------------------------------------------
// One Stockham pass: twiddle, E/R in-register DFTs, scatter into shared memory.
// ---------------------------------------------------------------------------------------
template <int N, int P, int R, int NS>
__device__ __forceinline__ void fft_pass_scatter(float2 (&v)[16], float2* __restrict__ s,
                                                 const float2* __restrict__ tw, int j, int p) {
    constexpr int NT = N / 16;
    constexpr int U = 16 / R;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int b = j + u * NT;
        const int k = b & (NS - 1);
        c2 x[R];
#pragma unroll
        for (int r = 0; r < R; ++r) x[r] = pk(v[u + r * U].x, v[u + r * U].y);
        if (NS > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) {
                const float2 w = __ldg(&tw[(r - 1) * NS + k]);
                x[r] = cmul2(x[r], w.x, w.y);
            }
        }
        RadixP<R>::run(x);
        const int j0 = (b - k) * R + k;
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const c2 y = x[Radix<R>::reg_of(q)];
            s[fft_pad(j0 + q * NS) * P + p] = make_float2(re(y), im(y));
        }
    }
}

template <int N, int P>
__device__ __forceinline__ void fft_gather(float2 (&v)[16], const float2* __restrict__ s, int j, int p) {
    constexpr int NT = N / 16;
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = s[fft_pad(j + NT * e) * P + p];
}

// number of twiddle words the pass with stride NS and radix R consumes
__host__ __device__ constexpr int fft_pass_tw_len(int R, int NS) { return NS > 1 ? (R - 1) * NS : 0; }

// Recursion over the radix-16 passes (I = 0..Q16-1), NS = 16^I
template <int LOG2N, int P, int I>
struct Pass16 {
    static constexpr int N = 1 << LOG2N;
    static constexpr int Q16 = LOG2N / 4;
    static constexpr int NS = 1 << (4 * I);
    static constexpr int RLAST = 1 << (LOG2N % 4);
    __device__ __forceinline__ static void run(float2 (&v)[16], float2* s, const float2* tw, int j, int p) {
        if constexpr (I < Q16) {
            fft_pass_scatter<N, P, 16, NS>(v, s, tw, j, p);
            constexpr bool last = (I == Q16 - 1) && (RLAST == 1);
            __syncthreads();
            if constexpr (!last) {
                fft_gather<N, P>(v, s, j, p);
                __syncthreads();
                Pass16<LOG2N, P, I + 1>::run(v, s, tw + fft_pass_tw_len(16, NS), j, p);
            }
        } else if constexpr (RLAST > 1) {
            fft_pass_scatter<N, P, RLAST, NS>(v, s, tw, j, p);
            __syncthreads();
        }
    }
};

// Forward complex FFT.  In: v[e] = element (j + NT*e).  Out: natural-order result in shared
// memory at s[fft_pad(k)*P + p], visible to the whole block (ends with __syncthreads()).
// All threads of the block must call this together.
template <int LOG2N, int P>
__device__ __forceinline__ void block_fft(float2 (&v)[16], float2* s, const float2* tw, int j, int p) {
    static_assert(LOG2N >= 4 && LOG2N <= 14, "block_fft supports 16 <= N <= 16384");
    Pass16<LOG2N, P, 0>::run(v, s, tw, j, p);
}

}  // namespace spyb
