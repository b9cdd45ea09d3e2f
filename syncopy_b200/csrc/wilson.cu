// K7-K9: regularisation, Wilson spectral factorisation and Geweke-Granger causality as batched
// per-frequency FP64 kernels.
//
// Replaces the NumPy / LAPACK bodies of
//   syncopy/connectivity/wilson_sf.py:197-254   regularize_csd   (2-norm condition numbers + eps ladder)
//   syncopy/connectivity/wilson_sf.py:16-194    wilson_sf, _psi0_initial, _plusOperator, max_rel_err
//   syncopy/connectivity/granger.py:10-79       granger
//
// Every matrix is a row-major [C][C] complex128 (double2) block of a [nFreq][C][C] stack.  The reference
// mirrors the spectrum to negative frequencies (wilson_sf.py:63) and carries 2(nFreq-1) matrices; everything
// it computes there is the element-wise conjugate of the positive half, so the kernels only ever hold the
// nFreq one-sided matrices and the mirror is applied where the frequency-axis FFT needs it.
//
//   zpotrf_kernel      blocked right-looking Cholesky, one CTA per frequency          (wilson_sf.py:76,147)
//   zgesv_kernel       blocked LU with partial pivoting on the augmented [psi | L], then blocked back
//                      substitution: X = psi^-1 L, one CTA per frequency               (wilson_sf.py:80-86)
//   zgemm_kernel       64x64x8 register-tiled complex GEMM; variants: A*B, A*B^H, upper tiles only, and an
//                      epilogue that reduces max |S - A A^H| / |S| instead of storing   (wilson_sf.py:87,101,105)
//   plus_* / zfft_*    the []+ operator (wilson_sf.py:154-184): frequency-axis FFTs as Stockham passes over
//                      [freq][matrix element] with one thread per element (coalesced, no shared memory)
//   zhetrd_kernel, tridiag_cond_kernel   Householder tridiagonalisation + Sturm bisection for the extreme
//                      singular values of the Hermitian CSD matrices                   (wilson_sf.py:239,248)
//   granger_kernel     granger.py:53-77 element-wise
#include "common.cuh"
#include "spyb_internal.h"

#include <cmath>
#include <vector>

namespace spyb {
namespace {

typedef double2 zd;

__device__ __forceinline__ zd zmk(double x, double y) { return make_double2(x, y); }
__device__ __forceinline__ zd zconj(zd a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ zd zscal(zd a, double s) { return make_double2(a.x * s, a.y * s); }
__device__ __forceinline__ void zfma(zd& c, zd a, zd b) {        // c += a * b
    c.x = fma(a.x, b.x, c.x); c.x = fma(-a.y, b.y, c.x);
    c.y = fma(a.x, b.y, c.y); c.y = fma(a.y, b.x, c.y);
}
__device__ __forceinline__ void zfms(zd& c, zd a, zd b) {        // c -= a * b
    c.x = fma(-a.x, b.x, c.x); c.x = fma(a.y, b.y, c.x);
    c.y = fma(-a.x, b.y, c.y); c.y = fma(-a.y, b.x, c.y);
}
__device__ __forceinline__ zd zrecip(zd b) {                     // 1 / b (Smith)
    if (fabs(b.x) >= fabs(b.y)) {
        const double r = b.y / b.x, d = b.x + b.y * r;
        return make_double2(1.0 / d, -r / d);
    }
    const double r = b.x / b.y, d = b.x * r + b.y;
    return make_double2(r / d, -1.0 / d);
}
__device__ __forceinline__ double zabs1(zd a) { return fabs(a.x) + fabs(a.y); }   // LAPACK cabs1 (izamax)

constexpr int NB = 16;            // panel width of the blocked factorizations
constexpr int PLD = NB + 1;       // padded panel row length (conflict-free column walks)
constexpr int FACT_THREADS = 256;
constexpr int MAX_CHAN = 256;     // largest matrix the factorization kernels stage in shared memory

// ------------------------------------------------------------------------------------------------------
// M[r][c] -= sum_q Ps[(r - prow0)][q] * Us[q][(c - ucol0)]   for r in [r0, r1), c in [c0, c1)
// Ps: panel rows (padded, [..][PLD]); Us: [NB][uld].  4 x 2 register tiles, lanes along columns.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tile_update(zd* __restrict__ M, int ld, int r0, int r1, int c0, int c1,
                                            const zd* __restrict__ Ps, int prow0,
                                            const zd* __restrict__ Us, int uld, int ucol0, int nbk) {
    const int nr = r1 - r0, nc = c1 - c0;
    if (nr <= 0 || nc <= 0) return;
    const int half = (nc + 1) >> 1;
    const int rb_n = (nr + 3) >> 2;
    const int items = rb_n * half;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int rb = it / half, cb = it - rb * half;
        const int ra = r0 + rb * 4;
        const int ca = c0 + cb, cbb = ca + half;
        const bool c2ok = cbb < c1;
        zd acc[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool rok = ra + i < r1;
            acc[i][0] = rok ? M[(long long)(ra + i) * ld + ca] : zmk(0, 0);
            acc[i][1] = (rok && c2ok) ? M[(long long)(ra + i) * ld + cbb] : zmk(0, 0);
        }
        const zd* pr = Ps + (ra - prow0) * PLD;
        const zd* u0 = Us + (ca - ucol0);
        const zd* u1 = Us + ((c2ok ? cbb : ca) - ucol0);
        for (int q = 0; q < nbk; ++q) {
            const zd ua = u0[q * uld], ub = u1[q * uld];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // rows past r1 read panel rows that exist in shared memory only if ra + i < r1
                const zd l = (ra + i < r1) ? pr[i * PLD + q] : zmk(0, 0);
                zfms(acc[i][0], l, ua);
                zfms(acc[i][1], l, ub);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (ra + i < r1) {
                M[(long long)(ra + i) * ld + ca] = acc[i][0];
                if (c2ok) M[(long long)(ra + i) * ld + cbb] = acc[i][1];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Cholesky A = L L^H (lower), in place on a copy; the strict upper triangle of the result is zeroed.
// info[b] = 1 when a non-positive pivot shows up (np.linalg.cholesky would raise LinAlgError).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FACT_THREADS) zpotrf_kernel(const zd* __restrict__ Ain, long long sAin,
                                                              zd* __restrict__ Lout, long long sL, int n,
                                                              int* __restrict__ info) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    zd* Ps = reinterpret_cast<zd*>(smem_raw);                 // [n][PLD]
    zd* Us = Ps + (size_t)n * PLD;                            // [NB][n]
    __shared__ int s_bad;
    const int b = blockIdx.x, tid = threadIdx.x;
    const zd* A0 = Ain + (long long)b * sAin;
    zd* L = Lout + (long long)b * sL;
    if (tid == 0) s_bad = 0;
    for (int i = tid; i < n * n; i += blockDim.x) L[i] = A0[i];
    __syncthreads();

    for (int k0 = 0; k0 < n; k0 += NB) {
        const int nbk = min(NB, n - k0), m = n - k0;
        for (int i = tid; i < m * nbk; i += blockDim.x) {
            const int r = i / nbk, c = i - r * nbk;
            Ps[r * PLD + c] = L[(long long)(k0 + r) * n + k0 + c];
        }
        __syncthreads();
        for (int c = 0; c < nbk; ++c) {
            const double d2 = Ps[c * PLD + c].x;
            if (!(d2 > 0.0)) { if (tid == 0) s_bad = 1; }
            const double d = sqrt(d2 > 0.0 ? d2 : 1.0), inv = 1.0 / d;
            __syncthreads();
            if (tid == 0) Ps[c * PLD + c] = zmk(d, 0.0);
            for (int r = c + 1 + tid; r < m; r += blockDim.x) Ps[r * PLD + c] = zscal(Ps[r * PLD + c], inv);
            __syncthreads();
            // remaining panel columns cc > c: Ps[r][cc] -= Ps[r][c] * conj(Ps[cc][c]), r >= cc
            for (int i = tid; i < (m - c - 1) * (nbk - c - 1); i += blockDim.x) {
                const int r = c + 1 + i / (nbk - c - 1), cc = c + 1 + i % (nbk - c - 1);
                if (r >= cc) {
                    zd v = Ps[r * PLD + cc];
                    zfms(v, Ps[r * PLD + c], zconj(Ps[cc * PLD + c]));
                    Ps[r * PLD + cc] = v;
                }
            }
            __syncthreads();
        }
        // write the panel back (zero above the diagonal) and stage conj(L21)^T for the trailing update
        for (int i = tid; i < m * nbk; i += blockDim.x) {
            const int r = i / nbk, c = i - r * nbk;
            L[(long long)(k0 + r) * n + k0 + c] = (r >= c) ? Ps[r * PLD + c] : zmk(0, 0);
            if (r >= nbk) Us[c * n + (r - nbk)] = zconj(Ps[r * PLD + c]);
        }
        __syncthreads();
        tile_update(L, n, k0 + nbk, n, k0 + nbk, n, Ps, k0, Us, n, k0 + nbk, nbk);
        __syncthreads();
    }
    // rows k0..: entries right of the panel belong to the strict upper triangle
    for (int i = tid; i < n * n; i += blockDim.x) {
        const int r = i / n, c = i - r * n;
        if (c > r) L[i] = zmk(0, 0);
    }
    if (tid == 0 && s_bad) info[b] = 1;
}

// ------------------------------------------------------------------------------------------------------
// X = A^-1 B: LU with partial pivoting on the augmented [A | B] (the elimination carries B along, so L is
// never stored), then blocked back substitution with U.  A is copied to `Aw` (destroyed), B to `X`.
// info[b] = 2 on an exactly zero pivot (np.linalg.inv: "Singular matrix").
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FACT_THREADS) zgesv_kernel(const zd* __restrict__ Ain, long long sAin,
                                                             const zd* __restrict__ Bin, long long sBin,
                                                             zd* __restrict__ Awork, long long sAw,
                                                             zd* __restrict__ Xout, long long sX, int n,
                                                             int* __restrict__ info) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    zd* Ps = reinterpret_cast<zd*>(smem_raw);                 // [n][PLD]
    zd* Us = Ps + (size_t)n * PLD;                            // [NB][2n]
    __shared__ double s_val[FACT_THREADS / 32];
    __shared__ int s_idx[FACT_THREADS / 32];
    __shared__ int s_piv[NB];
    __shared__ zd s_rdiag[NB];
    __shared__ int s_bad;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const zd* A0 = Ain + (long long)b * sAin;
    const zd* B0 = Bin + (long long)b * sBin;
    zd* A = Awork + (long long)b * sAw;
    zd* X = Xout + (long long)b * sX;
    const int w2 = 2 * n;
    if (tid == 0) s_bad = 0;
    for (int i = tid; i < n * n; i += blockDim.x) { A[i] = A0[i]; X[i] = B0[i]; }
    __syncthreads();

    for (int k0 = 0; k0 < n; k0 += NB) {
        const int nbk = min(NB, n - k0), m = n - k0;
        for (int i = tid; i < m * nbk; i += blockDim.x) {
            const int r = i / nbk, c = i - r * nbk;
            Ps[r * PLD + c] = A[(long long)(k0 + r) * n + k0 + c];
        }
        __syncthreads();
        // ---- panel factorization with partial pivoting ----
        for (int c = 0; c < nbk; ++c) {
            double best = -1.0; int bi = c;
            for (int r = c + tid; r < m; r += blockDim.x) {
                const double v = zabs1(Ps[r * PLD + c]);
                if (v > best) { best = v; bi = r; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, best, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
            __syncthreads();
            if (tid == 0) {
                double bb = s_val[0]; int ii = s_idx[0];
                for (int w = 1; w < FACT_THREADS / 32; ++w)
                    if (s_val[w] > bb || (s_val[w] == bb && s_idx[w] < ii)) { bb = s_val[w]; ii = s_idx[w]; }
                s_piv[c] = ii;
                if (!(bb > 0.0)) s_bad = 1;
            }
            __syncthreads();
            const int pv = s_piv[c];
            if (pv != c && tid < nbk) {
                const zd t = Ps[c * PLD + tid];
                Ps[c * PLD + tid] = Ps[pv * PLD + tid];
                Ps[pv * PLD + tid] = t;
            }
            __syncthreads();
            const zd pinv = zrecip(Ps[c * PLD + c]);
            for (int r = c + 1 + tid; r < m; r += blockDim.x) {
                const zd l = cmul(Ps[r * PLD + c], pinv);
                Ps[r * PLD + c] = l;
                for (int cc = c + 1; cc < nbk; ++cc) {
                    zd v = Ps[r * PLD + cc];
                    zfms(v, l, Ps[c * PLD + cc]);
                    Ps[r * PLD + cc] = v;
                }
            }
            __syncthreads();
        }
        // ---- U11 back to A; row swaps + forward substitution for every column right of the panel ----
        for (int i = tid; i < nbk * nbk; i += blockDim.x) {
            const int r = i / nbk, c = i - r * nbk;
            if (c >= r) A[(long long)(k0 + r) * n + k0 + c] = Ps[r * PLD + c];
        }
        const int cfirst = k0 + nbk;                              // augmented columns [cfirst, 2n)
        for (int col = cfirst + tid; col < w2; col += blockDim.x) {
            zd* base = col < n ? (A + col) : (X + (col - n));
            zd v[NB];
            for (int c = 0; c < nbk; ++c) {
                const int pv = s_piv[c];
                if (pv != c) {
                    const zd t = base[(long long)(k0 + c) * n];
                    base[(long long)(k0 + c) * n] = base[(long long)(k0 + pv) * n];
                    base[(long long)(k0 + pv) * n] = t;
                }
            }
#pragma unroll
            for (int r = 0; r < NB; ++r) v[r] = r < nbk ? base[(long long)(k0 + r) * n] : zmk(0, 0);
#pragma unroll
            for (int r = 1; r < NB; ++r) {
                if (r < nbk) {
#pragma unroll
                    for (int q = 0; q < r; ++q) zfms(v[r], Ps[r * PLD + q], v[q]);
                }
            }
#pragma unroll
            for (int r = 0; r < NB; ++r) {
                if (r < nbk) {
                    base[(long long)(k0 + r) * n] = v[r];
                    Us[r * w2 + (col - cfirst)] = v[r];
                }
            }
        }
        __syncthreads();
        // ---- trailing update of A (columns right of the panel) and of the carried right-hand sides ----
        tile_update(A, n, k0 + nbk, n, cfirst, n, Ps, k0, Us, w2, cfirst, nbk);
        tile_update(X, n, k0 + nbk, n, 0, n, Ps, k0, Us, w2, cfirst - n, nbk);
        __syncthreads();
    }

    // ---- back substitution U X = Y, bottom block row first ----
    const int last = ((n - 1) / NB) * NB;
    for (int k0 = last; k0 >= 0; k0 -= NB) {
        const int nbk = min(NB, n - k0);
        // rows [0, k0 + nbk) of the block column: U12 (rows above) and U11
        for (int i = tid; i < (k0 + nbk) * nbk; i += blockDim.x) {
            const int r = i / nbk, c = i - r * nbk;
            Ps[r * PLD + c] = A[(long long)r * n + k0 + c];
        }
        __syncthreads();
        if (tid < nbk) s_rdiag[tid] = zrecip(Ps[(k0 + tid) * PLD + tid]);
        __syncthreads();
        for (int col = tid; col < n; col += blockDim.x) {
            zd v[NB];
#pragma unroll
            for (int r = 0; r < NB; ++r) v[r] = r < nbk ? X[(long long)(k0 + r) * n + col] : zmk(0, 0);
#pragma unroll
            for (int r = NB - 1; r >= 0; --r) {
                if (r < nbk) {
#pragma unroll
                    for (int q = r + 1; q < NB; ++q)
                        if (q < nbk) zfms(v[r], Ps[(k0 + r) * PLD + q], v[q]);
                    v[r] = cmul(v[r], s_rdiag[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < NB; ++r) {
                if (r < nbk) {
                    X[(long long)(k0 + r) * n + col] = v[r];
                    Us[r * w2 + col] = v[r];
                }
            }
        }
        __syncthreads();
        tile_update(X, n, 0, k0, 0, n, Ps, 0, Us, w2, 0, nbk);
        __syncthreads();
    }
    if (tid == 0 && s_bad) info[b] = 2;
}

// ------------------------------------------------------------------------------------------------------
// Batched complex GEMM  C[b] = A[b] * op(B[b]),  op = identity (OPB 0) or conjugate transpose (OPB 1).
// EPI 0 stores C; EPI 1 reduces max |S - C| / |S| (wilson_sf.py:190-194) into *err_bits instead.
// ------------------------------------------------------------------------------------------------------
struct GemmArgs {
    const zd* A; long long sA;
    const zd* B; long long sB;
    zd* C; long long sC;
    const zd* S; long long sS;
    unsigned long long* err_bits;
    int n;
    int upper_only;
};

template <int OPB, int EPI>
__global__ void __launch_bounds__(256) zgemm_kernel(const GemmArgs a) {
    constexpr int TM = 64, TN = 64, BK = 8;
    __shared__ zd As[BK][TM + 1];
    __shared__ zd Bs[BK][TN + 1];
    __shared__ double s_red[8];
    const int bx = blockIdx.x, by = blockIdx.y, b = blockIdx.z;
    if (a.upper_only && bx < by) return;
    const int n = a.n;
    const zd* __restrict__ A = a.A + (long long)b * a.sA;
    const zd* __restrict__ B = a.B + (long long)b * a.sB;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = by * TM, col0 = bx * TN;
    zd acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = zmk(0, 0);

    for (int k0 = 0; k0 < n; k0 += BK) {
        // A tile: 64 rows x 8 k, k fastest in memory
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * 256, kk = idx & 7, r = idx >> 3;
            const int gr = row0 + r, gk = k0 + kk;
            As[kk][r] = (gr < n && gk < n) ? A[(long long)gr * n + gk] : zmk(0, 0);
        }
        if (OPB == 0) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int idx = tid + i * 256, cc = idx & 63, kk = idx >> 6;
                const int gc = col0 + cc, gk = k0 + kk;
                Bs[kk][cc] = (gc < n && gk < n) ? B[(long long)gk * n + gc] : zmk(0, 0);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int idx = tid + i * 256, kk = idx & 7, cc = idx >> 3;
                const int gc = col0 + cc, gk = k0 + kk;
                Bs[kk][cc] = (gc < n && gk < n) ? zconj(B[(long long)gc * n + gk]) : zmk(0, 0);
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            zd av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[kk][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) zfma(acc[i][j], av[i], bv[j]);
        }
        __syncthreads();
    }

    if (EPI == 0) {
        zd* __restrict__ C = a.C + (long long)b * a.sC;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int gr = row0 + ty + 16 * i;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gc = col0 + tx + 16 * j;
                if (gr < n && gc < n) C[(long long)gr * n + gc] = acc[i][j];
            }
        }
    } else {
        const zd* __restrict__ S = a.S + (long long)b * a.sS;
        double e = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int gr = row0 + ty + 16 * i;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gc = col0 + tx + 16 * j;
                if (gr < n && gc < n) {
                    const zd s = S[(long long)gr * n + gc];
                    const double num = hypot(s.x - acc[i][j].x, s.y - acc[i][j].y);
                    const double den = hypot(s.x, s.y);
                    const double q = num / den;
                    // NaN (0/0) and inf must not be lost: they mean "not converged"
                    e = (q == q) ? fmax(e, q) : INFINITY;
                }
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) e = fmax(e, __shfl_xor_sync(0xffffffffu, e, off));
        if ((tid & 31) == 0) s_red[tid >> 5] = e;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 8; ++w) e = fmax(e, s_red[w]);
            atomicMax(a.err_bits, (unsigned long long)__double_as_longlong(e));   // non-negative doubles order as integers
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Frequency-axis FFT: data [len][E] complex128, one thread per element column, Stockham passes.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void zdft2(zd& a, zd& b) { const zd t = a; a = cadd(t, b); b = csub(t, b); }
__device__ __forceinline__ void zdft4(zd& a0, zd& a1, zd& a2, zd& a3) {
    const zd s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
    a0 = cadd(s02, s13);
    a2 = csub(s02, s13);
    a1 = zmk(d02.x + d13.y, d02.y - d13.x);
    a3 = zmk(d02.x - d13.y, d02.y + d13.x);
}

// in-register forward DFT of R points; the result X[k] is left in x[k]
template <int R>
__device__ __forceinline__ void zdft(zd (&x)[R], const zd* __restrict__ tw, int len) {
    if constexpr (R == 2) {
        zdft2(x[0], x[1]);
    } else if constexpr (R == 4) {
        zdft4(x[0], x[1], x[2], x[3]);
    } else if constexpr (R == 16) {
        // n = i + 4m, k = q + 4s: T_i[q] in x[i + 4q]; X[q + 4s] in x[s + 4q]
        zdft4(x[0], x[4], x[8], x[12]);
        zdft4(x[1], x[5], x[9], x[13]);
        zdft4(x[2], x[6], x[10], x[14]);
        zdft4(x[3], x[7], x[11], x[15]);
        const double h = 0.70710678118654752440, c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;
        x[5] = cmul(x[5], zmk(c1, -s1));
        { const zd t = x[9]; x[9] = zmk((t.x + t.y) * h, (t.y - t.x) * h); }
        x[13] = cmul(x[13], zmk(s1, -c1));
        { const zd t = x[6]; x[6] = zmk((t.x + t.y) * h, (t.y - t.x) * h); }
        x[10] = zmk(x[10].y, -x[10].x);
        { const zd t = x[14]; x[14] = zmk((t.y - t.x) * h, -(t.x + t.y) * h); }
        x[7] = cmul(x[7], zmk(s1, -c1));
        { const zd t = x[11]; x[11] = zmk((t.y - t.x) * h, -(t.x + t.y) * h); }
        x[15] = cmul(x[15], zmk(-c1, s1));
        zdft4(x[0], x[1], x[2], x[3]);
        zdft4(x[4], x[5], x[6], x[7]);
        zdft4(x[8], x[9], x[10], x[11]);
        zdft4(x[12], x[13], x[14], x[15]);
        // un-permute: X[q + 4s] sits in x[s + 4q]
        zd y[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) y[k] = x[(k >> 2) + 4 * (k & 3)];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = y[k];
    } else {
        // small odd (or 8-point) DFT straight from the twiddle table: W_R^m = tw[m * len / R]
        zd y[R];
        const int step = len / R;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            zd s = x[0];
#pragma unroll
            for (int t = 1; t < R; ++t) zfma(s, x[t], tw[((k * t) % R) * step]);
            y[k] = s;
        }
#pragma unroll
        for (int k = 0; k < R; ++k) x[k] = y[k];
    }
}

template <int R>
__global__ void __launch_bounds__(128) zfft_pass_kernel(const zd* __restrict__ in, zd* __restrict__ out,
                                                        const zd* __restrict__ tw, int len, int ns,
                                                        long long E) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int j = blockIdx.y;                 // butterfly index, 0 .. len/R - 1
    const int k = j % ns;
    const int stride = len / R;
    const int tstep = len / (ns * R);
    zd x[R];
#pragma unroll
    for (int t = 0; t < R; ++t) x[t] = in[((long long)j + (long long)t * stride) * E + e];
    if (ns > 1) {
#pragma unroll
        for (int t = 1; t < R; ++t) x[t] = cmul(x[t], tw[(long long)t * k * tstep]);
    }
    zdft<R>(x, tw, len);
    const long long j0 = (long long)(j - k) * R + k;
#pragma unroll
    for (int q = 0; q < R; ++q) out[(j0 + (long long)q * ns) * E + e] = x[q];
}

// any radix (large prime factors): one output per thread, O(R) inputs each
__global__ void __launch_bounds__(128) zfft_pass_generic_kernel(const zd* __restrict__ in, zd* __restrict__ out,
                                                                const zd* __restrict__ tw, int len, int ns,
                                                                int R, long long E) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int jq = blockIdx.y;                // j * R + q
    const int j = jq / R, q = jq - j * R;
    const int k = j % ns;
    const int stride = len / R;
    const int tstep = len / (ns * R);
    zd s = zmk(0, 0);
    for (int t = 0; t < R; ++t) {
        const zd v = in[((long long)j + (long long)t * stride) * E + e];
        const zd w1 = tw[(long long)t * k * tstep];
        const zd w2 = tw[(long long)(((long long)q * t) % R) * stride];
        zfma(s, cmul(v, w1), w2);
    }
    const long long j0 = (long long)(j - k) * R + k;
    out[(j0 + (long long)q * ns) * E + e] = s;
}

__global__ void twiddle_kernel(zd* tw, int len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    double s, c;
    sincospi(-2.0 * (double)i / (double)len, &s, &c);
    tw[i] = zmk(c, s);
}

// upper-triangle element table: e -> (i, j), i <= j, row-major
__global__ void pair_table_kernel(int2* tab, int n) {
    const int i = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || j >= n || j < i) return;
    const long long e = (long long)i * n - (long long)i * (i - 1) / 2 + (j - i);
    tab[e] = make_int2(i, j);
}

// W[f][e] = conj(G[f][i][j] + delta_ij) mirrored to the full circle (forward FFT of the conjugate = N * ifft,
// conjugated; only its real part is used afterwards)
__global__ void __launch_bounds__(128) plus_pack_kernel(const zd* __restrict__ g, int nf, int n,
                                                        const int2* __restrict__ tab, long long E,
                                                        zd* __restrict__ W, int row0) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int f = row0 + blockIdx.y;          // row of the full circle, 0 .. len-1
    const int len = 2 * (nf - 1);
    const int2 ij = tab[e];
    const int fs = f < nf ? f : len - f;
    zd v = g[((long long)fs * n + ij.x) * n + ij.y];
    if (ij.x == ij.y) v.x += 1.0;
    // mirrored half is conj(G); W holds conj of the mirrored spectrum
    W[(long long)f * E + e] = f < nf ? zconj(v) : v;
}

// lag domain: beta = Re(W) / len; causal projection (wilson_sf.py:171-180) packed as
// Z[l] = w_l (beta_ij[l] + i beta_ij[len - l]) so that one forward FFT yields gplus_ij and gplus_ji.
// Also emits M0 = gplus_0 + S with S = triu(g0) - triu(g0)^H (wilson_sf.py:96-98): upper triangular.
__global__ void __launch_bounds__(128) plus_causal_kernel(const zd* __restrict__ W, int nf, int n,
                                                          const int2* __restrict__ tab, long long E,
                                                          zd* __restrict__ Z, zd* __restrict__ M0,
                                                          double* __restrict__ g0) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int l = blockIdx.y;
    const int len = 2 * (nf - 1), nlag = len / 2;
    zd z = zmk(0, 0);
    if (l <= nlag) {
        const double inv = 1.0 / (double)len;
        const double wgt = (l == 0 || l == nlag) ? 0.5 : 1.0;
        const double a = W[(long long)l * E + e].x * inv;
        const double bb = W[(long long)((len - l) % len) * E + e].x * inv;
        z = zmk(wgt * a, wgt * bb);
        if (l == 0) {
            const int2 ij = tab[e];
            const double v = 0.5 * a;
            g0[e] = v;
            if (ij.x == ij.y) {
                M0[(long long)ij.x * n + ij.y] = zmk(v, 0.0);
            } else {
                M0[(long long)ij.x * n + ij.y] = zmk(2.0 * v, 0.0);
                M0[(long long)ij.y * n + ij.x] = zmk(0.0, 0.0);
            }
        }
    }
    Z[(long long)l * E + e] = z;
}

// M[f] = gplus[f] + S for the one-sided frequencies (wilson_sf.py:101)
__global__ void __launch_bounds__(128) plus_unpack_kernel(const zd* __restrict__ Zf, int nf, int n,
                                                          const int2* __restrict__ tab, long long E,
                                                          const double* __restrict__ g0, zd* __restrict__ M,
                                                          int row0) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int f = row0 + blockIdx.y;          // 0 .. nf-1
    const int len = 2 * (nf - 1);
    const int2 ij = tab[e];
    const zd z1 = Zf[(long long)f * E + e];
    const zd z2 = zconj(Zf[(long long)((len - f) % len) * E + e]);
    const zd gij = zmk(0.5 * (z1.x + z2.x), 0.5 * (z1.y + z2.y));
    zd* Mf = M + (long long)f * n * n;
    if (ij.x == ij.y) {
        Mf[(long long)ij.x * n + ij.y] = gij;
    } else {
        const zd d = csub(z1, z2);
        const zd gji = zmk(0.5 * d.y, -0.5 * d.x);      // (z1 - z2) / (2i)
        const double s = g0[e];
        Mf[(long long)ij.x * n + ij.y] = zmk(gij.x + s, gij.y);
        Mf[(long long)ij.y * n + ij.x] = zmk(gji.x - s, gji.y);
    }
}

// gamma0 = Re( sym( sum over the mirrored circle of S ) ) as a complex matrix (wilson_sf.py:135-141)
__global__ void gamma0_kernel(const zd* __restrict__ S, int nf, int n, zd* __restrict__ G0) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    const int i = idx / n, j = idx - i * n;
    const long long n2 = (long long)n * n;
    double s1 = 0.0, s2 = 0.0;
    for (int f = 0; f < nf; ++f) {
        const double wgt = (f == 0 || f == nf - 1) ? 1.0 : 2.0;
        s1 += wgt * S[f * n2 + (long long)i * n + j].x;
        s2 += wgt * S[f * n2 + (long long)j * n + i].x;
    }
    G0[idx] = zmk(0.5 * (s1 + s2), 0.0);
}

// psi[f] = L0^T for every f, psi0 = L0^T (wilson_sf.py:66-67,151)
__global__ void psi_init_kernel(const zd* __restrict__ L0, int n, zd* __restrict__ psi0, zd* __restrict__ psi,
                                int f0) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    const int i = idx / n, j = idx - i * n;
    const zd v = L0[(long long)j * n + i];
    if (blockIdx.y == 0) psi0[idx] = v;
    psi[(long long)(f0 + blockIdx.y) * n * n + idx] = v;
}

__global__ void eye_kernel(zd* I, int n) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    I[idx] = zmk((idx / n) == (idx % n) ? 1.0 : 0.0, 0.0);
}

__global__ void real_part_kernel(const zd* __restrict__ src, double* __restrict__ dst, long long nelem) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nelem) dst[i] = src[i].x;
}

// ------------------------------------------------------------------------------------------------------
// regularize_csd: Hermitian tridiagonalisation (unblocked Householder, lower storage conventions of zhetd2)
// in FP64 on a complex128 copy, then Sturm-count bisection on the real symmetric tridiagonal.
// d, e are stored transposed ([i][freq]) so that the bisection kernel (one thread per frequency) coalesces.
// ------------------------------------------------------------------------------------------------------
constexpr int HT_THREADS = 256;

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < HT_THREADS / 32; ++w) s += red[w];
    return s;
}

__global__ void __launch_bounds__(HT_THREADS) zhetrd_kernel(const float2* __restrict__ csd, int n, int nf,
                                                            zd* __restrict__ work, double* __restrict__ dd,
                                                            double* __restrict__ ee) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    zd* v = reinterpret_cast<zd*>(smem_raw);     // [n]
    zd* p = v + n;                               // [n]
    __shared__ double red[HT_THREADS / 32];
    __shared__ zd s_tau;
    const int b = blockIdx.x, tid = threadIdx.x;
    const float2* A0 = csd + (long long)b * n * n;
    zd* A = work + (long long)b * n * n;
    // Hermitian part of the input (the reference feeds the matrix as is to an SVD; CSDs are Hermitian)
    for (int i = tid; i < n * n; i += blockDim.x) {
        const int r = i / n, c = i - r * n;
        const float2 x = A0[i], y = A0[(long long)c * n + r];
        A[i] = zmk(0.5 * ((double)x.x + (double)y.x), 0.5 * ((double)x.y - (double)y.y));
    }
    __syncthreads();
    for (int k = 0; k < n - 1; ++k) {
        const int m = n - k - 1;                 // length of the Householder vector (rows k+1 .. n-1)
        // column k below the diagonal
        double part = 0.0;
        for (int i = 1 + tid; i < m; i += blockDim.x) {
            const zd x = A[(long long)(k + 1 + i) * n + k];
            part += x.x * x.x + x.y * x.y;
        }
        const double xnorm2 = block_sum(part, red);
        const zd alpha = A[(long long)(k + 1) * n + k];
        double beta;
        zd tau;
        if (xnorm2 == 0.0 && alpha.y == 0.0) {
            tau = zmk(0, 0);
            beta = alpha.x;
        } else {
            const double nrm = sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xnorm2);
            beta = alpha.x >= 0.0 ? -nrm : nrm;
            tau = zmk((beta - alpha.x) / beta, -alpha.y / beta);
        }
        const zd sc = (tau.x == 0.0 && tau.y == 0.0) ? zmk(0, 0) : zrecip(zmk(alpha.x - beta, alpha.y));
        for (int i = tid; i < m; i += blockDim.x)
            v[i] = i == 0 ? zmk(1.0, 0.0) : cmul(A[(long long)(k + 1 + i) * n + k], sc);
        if (tid == 0) {
            dd[(long long)k * nf + b] = A[(long long)k * n + k].x;
            ee[(long long)k * nf + b] = beta;
            s_tau = tau;
        }
        __syncthreads();
        if (tau.x != 0.0 || tau.y != 0.0) {
            zd* A22 = A + (long long)(k + 1) * n + (k + 1);
            // p = tau * A22 v, with (A22 v)_i = sum_j conj(A22[j][i]) v_j  (Hermitian: coalesced column walk)
            for (int i = tid; i < m; i += blockDim.x) {
                zd s = zmk(0, 0);
                for (int j = 0; j < m; ++j) zfma(s, zconj(A22[(long long)j * n + i]), v[j]);
                p[i] = cmul(tau, s);
            }
            __syncthreads();
            // w = p - (tau/2) (p^H v) v
            double pr = 0.0, pi = 0.0;
            for (int i = tid; i < m; i += blockDim.x) {
                const zd t = cmul(zconj(p[i]), v[i]);
                pr += t.x; pi += t.y;
            }
            const double dr = block_sum(pr, red);
            const double di = block_sum(pi, red);
            const zd al = cmul(zmk(-0.5 * tau.x, -0.5 * tau.y), zmk(dr, di));
            __syncthreads();
            for (int i = tid; i < m; i += blockDim.x) { zd w = p[i]; zfma(w, al, v[i]); p[i] = w; }
            __syncthreads();
            // A22 -= v w^H + w v^H
            for (int idx = tid; idx < m * m; idx += blockDim.x) {
                const int i = idx / m, j = idx - i * m;
                zd x = A22[(long long)i * n + j];
                zfms(x, v[i], zconj(p[j]));
                zfms(x, p[i], zconj(v[j]));
                A22[(long long)i * n + j] = x;
            }
        }
        __syncthreads();
    }
    if (tid == 0) dd[(long long)(n - 1) * nf + b] = A[(long long)(n - 1) * n + (n - 1)].x;
}

// number of eigenvalues of the tridiagonal (d + shift, e) that are < x
__device__ __forceinline__ int sturm_count(const double* __restrict__ d, const double* __restrict__ e, int n,
                                           long long stride, double shift, double x, double tiny) {
    int cnt = 0;
    double q = d[0] + shift - x;
    if (q < 0.0) ++cnt;
    for (int i = 1; i < n; ++i) {
        if (fabs(q) < tiny) q = q < 0.0 ? -tiny : tiny;
        const double ei = e[(long long)(i - 1) * stride];
        q = d[(long long)i * stride] + shift - x - ei * ei / q;
        if (q < 0.0) ++cnt;
    }
    return cnt;
}

__device__ double kth_eigenvalue(const double* d, const double* e, int n, long long stride, double shift,
                                 int k, double lo, double hi, double tiny) {
    for (int it = 0; it < 120; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (mid <= lo || mid >= hi) break;
        if (sturm_count(d, e, n, stride, shift, mid, tiny) > k) hi = mid; else lo = mid;
    }
    return 0.5 * (lo + hi);
}

// cond[f] = sigma_max / sigma_min of (A_f + shift I), sigma = |eigenvalue|
__global__ void tridiag_cond_kernel(const double* __restrict__ dd, const double* __restrict__ ee, int n, int nf,
                                    double shift, double* __restrict__ cond) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const double* d = dd + f;
    const double* e = ee + f;
    // Gershgorin interval
    double lo = INFINITY, hi = -INFINITY, scale = 0.0;
    for (int i = 0; i < n; ++i) {
        const double c = d[(long long)i * nf] + shift;
        const double r = (i > 0 ? fabs(e[(long long)(i - 1) * nf]) : 0.0) + (i < n - 1 ? fabs(e[(long long)i * nf]) : 0.0);
        lo = fmin(lo, c - r); hi = fmax(hi, c + r);
        scale = fmax(scale, fabs(c) + r);
    }
    const double tiny = fmax(scale, 1e-300) * 1e-300 + 1e-306;
    const double pad = 1e-12 * scale + 1e-300;
    lo -= pad; hi += pad;
    const double l0 = kth_eigenvalue(d, e, n, nf, shift, 0, lo, hi, tiny);
    const double ln = kth_eigenvalue(d, e, n, nf, shift, n - 1, lo, hi, tiny);
    const double smax = fmax(fabs(l0), fabs(ln));
    const int neg = sturm_count(d, e, n, nf, shift, 0.0, tiny);       // eigenvalues < 0
    double smin;
    if (neg == 0) smin = fabs(l0);
    else if (neg == n) smin = fabs(ln);
    else {
        const double a = kth_eigenvalue(d, e, n, nf, shift, neg - 1, lo, 0.0, tiny);   // largest negative
        const double b = kth_eigenvalue(d, e, n, nf, shift, neg, 0.0, hi, tiny);       // smallest non-negative
        smin = fmin(fabs(a), fabs(b));
    }
    cond[f] = smax / smin;          // inf for an exactly singular matrix, like np.linalg.cond
}

__global__ void max_reduce_kernel(const double* __restrict__ x, int n, double* __restrict__ out) {
    __shared__ double red[32];
    double v = 0.0;
    bool nan = false;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double t = x[i];
        if (t != t) nan = true; else v = fmax(v, t);
    }
    if (nan) v = INFINITY;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) v = fmax(v, red[w]);
        *out = v;
    }
}

// out = complex128(csd) + eps * I   (wilson_sf.py:247; eps = 0 is the plain cast of AV_compRoutines.py:395)
__global__ void regularized_copy_kernel(const float2* __restrict__ csd, long long total, int n, double eps,
                                        zd* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long within = i % ((long long)n * n);
    const int r = (int)(within / n), c = (int)(within - (long long)r * n);
    const float2 v = csd[i];
    out[i] = zmk((double)v.x + (r == c ? eps : 0.0), (double)v.y);
}

// granger.py:53-77
__global__ void granger_kernel(const zd* __restrict__ csd, const zd* __restrict__ H, const double* __restrict__ Sigma,
                               int nf, int n, float* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n2 = (long long)n * n;
    if (idx >= nf * n2) return;
    const long long f = idx / n2;
    const int i = (int)((idx - f * n2) / n), j = (int)(idx - f * n2 - (long long)i * n);
    const zd sj = csd[f * n2 + (long long)j * n + j];
    const double Sjj = hypot(sj.x, sj.y);
    const zd h = H[f * n2 + (long long)j * n + i];
    const double h2 = h.x * h.x + h.y * h.y;
    const double sig_ii = fabs(Sigma[(long long)i * n + i]), sig_jj = fabs(Sigma[(long long)j * n + j]);
    const double sig_ji = fabs(Sigma[(long long)j * n + i]);
    const double fac = sig_ii - sig_ji * sig_ji / sig_jj;
    out[idx] = (float)log(Sjj / (Sjj - fac * h2));
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
struct FftPlanZ {
    std::vector<int> radices;
};

FftPlanZ plan_fft(int len) {
    FftPlanZ p;
    int r = len;
    const int pref[] = {16, 8, 4, 2, 3, 5, 7};
    for (int f : pref)
        while (r % f == 0 && r > 1) { p.radices.push_back(f); r /= f; }
    for (int f = 11; r > 1; f += 2) {
        while (r % f == 0) { p.radices.push_back(f); r /= f; }
        if ((long long)f * f > r && r > 1) { p.radices.push_back(r); r = 1; }
    }
    return p;
}

template <int R>
int launch_pass(const zd* in, zd* out, const zd* tw, int len, int ns, long long E, cudaStream_t st) {
    if (len / R > 65535) return fail("frequency-axis FFT: length %d is too long", len);
    dim3 grid((unsigned)((E + 127) / 128), (unsigned)(len / R));
    zfft_pass_kernel<R><<<grid, 128, 0, st>>>(in, out, tw, len, ns, E);
    SPYB_LAUNCH_CHECK("zfft_pass_kernel");
    count_launch();
    return 0;
}

// forward FFT along axis 0 of [len][E]; data starts in `a`, result pointer returned through *res (a or b)
int fft_axis0(zd* a, zd* b, const zd* tw, int len, long long E, const FftPlanZ& plan, zd** res, cudaStream_t st) {
    zd* src = a;
    zd* dst = b;
    int ns = 1;
    for (int R : plan.radices) {
        int rc;
        switch (R) {
            case 16: rc = launch_pass<16>(src, dst, tw, len, ns, E, st); break;
            case 8:  rc = launch_pass<8>(src, dst, tw, len, ns, E, st); break;
            case 4:  rc = launch_pass<4>(src, dst, tw, len, ns, E, st); break;
            case 2:  rc = launch_pass<2>(src, dst, tw, len, ns, E, st); break;
            case 3:  rc = launch_pass<3>(src, dst, tw, len, ns, E, st); break;
            case 5:  rc = launch_pass<5>(src, dst, tw, len, ns, E, st); break;
            case 7:  rc = launch_pass<7>(src, dst, tw, len, ns, E, st); break;
            default: {
                if (len > 65535) return fail("frequency-axis FFT: length %d with prime factor %d is not supported", len, R);
                dim3 grid((unsigned)((E + 127) / 128), (unsigned)len);
                zfft_pass_generic_kernel<<<grid, 128, 0, st>>>(src, dst, tw, len, ns, R, E);
                SPYB_LAUNCH_CHECK("zfft_pass_generic_kernel");
                count_launch();
                rc = 0;
            }
        }
        if (rc) return rc;
        ns *= R;
        zd* t = src; src = dst; dst = t;
    }
    *res = src;
    return 0;
}

size_t fact_smem(int n) { return ((size_t)n * PLD + (size_t)NB * 2 * n) * sizeof(zd); }

int ensure_fact_smem() {
    // per device / context attribute: set on every call (several engines may live in one process)
    const int mx = (int)fact_smem(MAX_CHAN);
    SPYB_CUDA(cudaFuncSetAttribute(zpotrf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    SPYB_CUDA(cudaFuncSetAttribute(zgesv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    return 0;
}

int run_potrf(const zd* A, long long sA, zd* L, long long sL, int n, int batch, int* info, cudaStream_t st) {
    zpotrf_kernel<<<batch, FACT_THREADS, fact_smem(n), st>>>(A, sA, L, sL, n, info);
    SPYB_LAUNCH_CHECK("zpotrf_kernel");
    count_launch();
    return 0;
}

int run_gesv(const zd* A, long long sA, const zd* B, long long sB, zd* Aw, long long sAw, zd* X, long long sX,
             int n, int batch, int* info, cudaStream_t st) {
    zgesv_kernel<<<batch, FACT_THREADS, fact_smem(n), st>>>(A, sA, B, sB, Aw, sAw, X, sX, n, info);
    SPYB_LAUNCH_CHECK("zgesv_kernel");
    count_launch();
    return 0;
}

template <int OPB, int EPI>
int run_gemm(const zd* A, long long sA, const zd* B, long long sB, zd* C, long long sC, const zd* S, long long sS,
             unsigned long long* err_bits, int n, int batch, bool upper, cudaStream_t st) {
    GemmArgs g;
    g.A = A; g.sA = sA; g.B = B; g.sB = sB; g.C = C; g.sC = sC; g.S = S; g.sS = sS;
    g.err_bits = err_bits; g.n = n; g.upper_only = upper ? 1 : 0;
    const int t = (n + 63) / 64;
    dim3 grid(t, t, batch);
    zgemm_kernel<OPB, EPI><<<grid, 256, 0, st>>>(g);
    SPYB_LAUNCH_CHECK("zgemm_kernel");
    count_launch();
    return 0;
}

struct Carver {
    unsigned char* base;
    size_t off = 0;
    explicit Carver(void* p) : base(static_cast<unsigned char*>(p)) {}
    template <typename T> T* take(size_t count) {
        off = (off + 255) & ~(size_t)255;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

struct WilsonBuffers {
    zd *Lchol, *psiA, *psiB, *Aw, *X, *fftA, *fftB, *G0, *L0, *psi0A, *psi0B, *M0, *eye, *inv0, *tmp0, *tw;
    int2* tab;
    double* g0;
    unsigned long long* err_bits;
    int* info;
    size_t bytes;
};

WilsonBuffers carve_wilson(void* work, int nf, int n) {
    WilsonBuffers w;
    Carver c(work);
    const size_t n2 = (size_t)n * n, big = (size_t)nf * n2;
    const size_t E = (size_t)n * (n + 1) / 2, len = 2 * (size_t)(nf - 1);
    w.Lchol = c.take<zd>(big); w.psiA = c.take<zd>(big); w.psiB = c.take<zd>(big);
    w.Aw = c.take<zd>(big); w.X = c.take<zd>(big);
    w.fftA = c.take<zd>(len * E); w.fftB = c.take<zd>(len * E);
    w.G0 = c.take<zd>(n2); w.L0 = c.take<zd>(n2); w.psi0A = c.take<zd>(n2); w.psi0B = c.take<zd>(n2);
    w.M0 = c.take<zd>(n2); w.eye = c.take<zd>(n2); w.inv0 = c.take<zd>(n2); w.tmp0 = c.take<zd>(n2);
    w.tw = c.take<zd>(len);
    w.tab = c.take<int2>(E);
    w.g0 = c.take<double>(E);
    w.err_bits = c.take<unsigned long long>(4);
    w.info = c.take<int>((size_t)nf + 8);
    w.bytes = c.off + 256;
    return w;
}

int check_info(int* info_dev, int count, const char* what, cudaStream_t st) {
    std::vector<int> h(count);
    SPYB_CUDA(cudaMemcpyAsync(h.data(), info_dev, sizeof(int) * count, cudaMemcpyDeviceToHost, st));
    SPYB_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < count; ++i) {
        if (h[i] == 1) return fail("%s: matrix %d is not positive definite (numpy.linalg.LinAlgError in the reference)", what, i);
        if (h[i] == 2) return fail("%s: matrix %d is singular (numpy.linalg.LinAlgError in the reference)", what, i);
    }
    return 0;
}

}  // namespace

long long wilson_workspace_bytes(int n_freq, int n_chan) {
    if (n_freq < 2 || n_chan < 1) return 0;
    return (long long)carve_wilson(nullptr, n_freq, n_chan).bytes;
}

long long regularize_workspace_bytes(int n_freq, int n_chan) {
    Carver c(nullptr);
    c.take<zd>((size_t)n_freq * n_chan * n_chan);
    c.take<double>((size_t)n_freq * n_chan);
    c.take<double>((size_t)n_freq * n_chan);
    c.take<double>((size_t)n_freq + 8);
    return (long long)c.off + 256;
}

int regularize_csd(const void* csd_c64, int n_freq, int n_chan, double cond_max, double eps_max, int n_steps,
                   void* out_c128, double* eps_host, double* cond0_host, void* work, long long work_bytes,
                   cudaStream_t st) {
    if (n_freq < 1 || n_chan < 1) return fail("regularize_csd: empty input");
    if (n_chan > 1024) return fail("regularize_csd: %d channels not supported (max 1024)", n_chan);
    if (work_bytes < regularize_workspace_bytes(n_freq, n_chan))
        return fail("regularize_csd: workspace too small (%lld < %lld bytes)", work_bytes,
                    regularize_workspace_bytes(n_freq, n_chan));
    Carver c(work);
    zd* A = c.take<zd>((size_t)n_freq * n_chan * n_chan);
    double* dd = c.take<double>((size_t)n_freq * n_chan);
    double* ee = c.take<double>((size_t)n_freq * n_chan);
    double* cond = c.take<double>((size_t)n_freq + 8);
    double* cmax = cond + n_freq;
    const float2* csd = static_cast<const float2*>(csd_c64);

    zhetrd_kernel<<<n_freq, HT_THREADS, 2 * (size_t)n_chan * sizeof(zd), st>>>(csd, n_chan, n_freq, A, dd, ee);
    SPYB_LAUNCH_CHECK("zhetrd_kernel");
    count_launch();

    auto max_cond = [&](double shift, double* result) -> int {
        tridiag_cond_kernel<<<(n_freq + 63) / 64, 64, 0, st>>>(dd, ee, n_chan, n_freq, shift, cond);
        SPYB_LAUNCH_CHECK("tridiag_cond_kernel");
        count_launch();
        max_reduce_kernel<<<1, 256, 0, st>>>(cond, n_freq, cmax);
        SPYB_LAUNCH_CHECK("max_reduce_kernel");
        count_launch();
        SPYB_CUDA(cudaMemcpyAsync(result, cmax, sizeof(double), cudaMemcpyDeviceToHost, st));
        SPYB_CUDA(cudaStreamSynchronize(st));
        return 0;
    };

    double c0 = 0.0;
    if (max_cond(0.0, &c0)) return 1;
    *cond0_host = c0;
    double eps_used = 0.0, eps_report = 0.0;
    if (!(c0 < cond_max)) {
        eps_report = -1.0;
        for (int s = 0; s < n_steps; ++s) {
            // np.logspace(-10, log10(eps_max), nSteps): y = s * step + start, the last point is `stop` itself
            const double stop = std::log10(eps_max);
            const double step = n_steps > 1 ? (stop + 10.0) / (double)(n_steps - 1) : 0.0;
            const double expo = (n_steps > 1 && s == n_steps - 1) ? stop : (double)s * step - 10.0;
            const double eps = std::pow(10.0, expo);
            eps_used = eps;
            double cm = 0.0;
            if (max_cond(eps, &cm)) return 1;
            if (cm < cond_max) { eps_report = eps; break; }
        }
    }
    *eps_host = eps_report;
    const long long total = (long long)n_freq * n_chan * n_chan;
    regularized_copy_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(csd, total, n_chan, eps_used,
                                                                             static_cast<zd*>(out_c128));
    SPYB_LAUNCH_CHECK("regularized_copy_kernel");
    count_launch();
    return 0;
}

// Frequency-slab sharding: everything in an iteration except the plus operator is independent per frequency, so a
// rank only factorises its slab [f_lo, f_hi).  The plus operator needs every frequency of every matrix element; it is
// replicated: each rank packs its rows of the lag-domain work array, `exchange(ctx, 0, ...)` lets the caller gather
// the rows of all ranks (stream-ordered, e.g. NCCL broadcasts), and `exchange(ctx, 1, ...)` max-reduces the error
// scalar.  With exchange == NULL (one rank) the slab is the whole axis.  H_out is written for the slab only.
int wilson_sf(const void* csd_c128, int n_freq, int n_chan, int n_iter, double rtol, void* H_out, double* Sigma_out,
              int* converged_host, double* err_host, int* iters_host, void* work, long long work_bytes,
              int f_lo, int f_hi, WilsonExchangeFn exchange, void* exchange_ctx, cudaStream_t st) {
    const int nf = n_freq, n = n_chan;
    if (nf < 2) return fail("wilson_sf needs at least two frequencies");
    if (n < 1 || n > MAX_CHAN) return fail("wilson_sf supports 1..%d channels (got %d)", MAX_CHAN, n);
    if (work_bytes < wilson_workspace_bytes(nf, n))
        return fail("wilson_sf: workspace too small (%lld < %lld bytes)", work_bytes, wilson_workspace_bytes(nf, n));
    if (exchange == nullptr) { f_lo = 0; f_hi = nf; }
    if (f_lo < 0 || f_hi > nf || f_lo > f_hi) return fail("wilson_sf: bad frequency slab [%d, %d) of %d", f_lo, f_hi, nf);
    if (ensure_fact_smem()) return 1;
    WilsonBuffers w = carve_wilson(work, nf, n);
    const zd* S = static_cast<const zd*>(csd_c128);
    const long long n2 = (long long)n * n;
    const long long E = (long long)n * (n + 1) / 2;
    const int len = 2 * (nf - 1);
    const FftPlanZ plan = plan_fft(len);
    const unsigned eb = (unsigned)((E + 127) / 128);
    const unsigned nb2 = (unsigned)((n2 + 255) / 256);
    const int nfl = f_hi - f_lo;                      // frequencies of this rank
    const long long so = (long long)f_lo * n2;        // element offset of the slab in every [nF][C][C] stack
    // rows of the full circle this rank packs: its slab and the mirror image of the slab's interior
    const int m_lo = f_lo > 1 ? f_lo : 1, m_hi = f_hi < nf - 1 ? f_hi : nf - 1;    // mirrored rows: len - f, f in [m_lo, m_hi)
    const int mrow0 = len - m_hi + 1, mrows = m_hi > m_lo ? m_hi - m_lo : 0;

    // A numerical failure (non-positive-definite CSD, singular psi: numpy.linalg.LinAlgError in the reference) is seen
    // only by the rank that owns the frequency.  All ranks must take the same decision, or the ones that carry on
    // would wait in the next collective for ever: the flag is max-reduced through the exchange callback.
    auto agree_failure = [&](int local_fail) -> int {
        if (!exchange) return local_fail;
        double v = local_fail ? 1.0 : 0.0;
        double* flag = reinterpret_cast<double*>(w.err_bits + 1);
        if (cudaMemcpyAsync(flag, &v, sizeof(v), cudaMemcpyHostToDevice, st) != cudaSuccess) return 1;
        if (cudaStreamSynchronize(st) != cudaSuccess) return 1;
        if (exchange(exchange_ctx, 1, flag, 8, 1)) return 1;
        if (cudaMemcpyAsync(&v, flag, sizeof(v), cudaMemcpyDeviceToHost, st) != cudaSuccess) return 1;
        if (cudaStreamSynchronize(st) != cudaSuccess) return 1;
        if (v != 0.0 && !local_fail)
            fail("wilson_sf: another rank found a non-positive-definite or singular matrix in its frequency slab "
                 "(numpy.linalg.LinAlgError in the reference)");
        return v != 0.0 ? 1 : 0;
    };

    SPYB_CUDA(cudaMemsetAsync(w.info, 0, sizeof(int) * ((size_t)nf + 8), st));
    twiddle_kernel<<<(len + 255) / 256, 256, 0, st>>>(w.tw, len);
    SPYB_LAUNCH_CHECK("twiddle_kernel"); count_launch();
    pair_table_kernel<<<dim3((n + 127) / 128, n), 128, 0, st>>>(w.tab, n);
    SPYB_LAUNCH_CHECK("pair_table_kernel"); count_launch();
    eye_kernel<<<nb2, 256, 0, st>>>(w.eye, n);
    SPYB_LAUNCH_CHECK("eye_kernel"); count_launch();

    // psi0 = chol(Re sym gamma_0)^T, psi = tile(psi0)     (wilson_sf.py:66-67,123-151); gamma_0 needs every frequency
    gamma0_kernel<<<nb2, 256, 0, st>>>(S, nf, n, w.G0);
    SPYB_LAUNCH_CHECK("gamma0_kernel"); count_launch();
    if (run_potrf(w.G0, 0, w.L0, 0, n, 1, w.info + nf, st)) return 1;
    // U = chol(CSD)                                         (wilson_sf.py:76)
    if (nfl > 0 && run_potrf(S + so, n2, w.Lchol + so, n2, n, nfl, w.info + f_lo, st)) return 1;
    if (agree_failure(check_info(w.info, nf + 1, "wilson_sf (Cholesky of the CSD)", st))) return 1;
    psi_init_kernel<<<dim3(nb2, nfl > 0 ? nfl : 1), 256, 0, st>>>(w.L0, n, w.psi0A, nfl > 0 ? w.psiA : w.tmp0 - so, f_lo);
    SPYB_LAUNCH_CHECK("psi_init_kernel"); count_launch();

    zd* psi = w.psiA;
    zd* psi_next = w.psiB;
    zd* psi0 = w.psi0A;
    zd* psi0_next = w.psi0B;
    bool converged = false;
    double err = INFINITY;
    int it = 0;
    for (it = 0; it < n_iter; ++it) {
        zd* g = w.Aw;
        if (nfl > 0) {
            // g = psi^-1 U (psi^-1 U)^H                      (wilson_sf.py:80-87)
            if (run_gesv(psi + so, n2, w.Lchol + so, n2, w.Aw + so, n2, w.X + so, n2, n, nfl, w.info + f_lo, st)) return 1;
            if (run_gemm<1, 0>(w.X + so, n2, w.X + so, n2, g + so, n2, nullptr, 0, nullptr, n, nfl, true, st)) return 1;
            // [g + I]+                                       (wilson_sf.py:94, 154-184)
            plus_pack_kernel<<<dim3(eb, nfl), 128, 0, st>>>(g, nf, n, w.tab, E, w.fftA, f_lo);
            SPYB_LAUNCH_CHECK("plus_pack_kernel"); count_launch();
            if (mrows > 0) {
                plus_pack_kernel<<<dim3(eb, mrows), 128, 0, st>>>(g, nf, n, w.tab, E, w.fftA, mrow0);
                SPYB_LAUNCH_CHECK("plus_pack_kernel"); count_launch();
            }
        }
        if (exchange && exchange(exchange_ctx, 0, w.fftA, E * (long long)sizeof(zd), len))
            return fail("wilson_sf: the exchange callback failed while gathering the packed spectra");
        zd* lag = nullptr;
        if (fft_axis0(w.fftA, w.fftB, w.tw, len, E, plan, &lag, st)) return 1;
        zd* other = lag == w.fftA ? w.fftB : w.fftA;
        plus_causal_kernel<<<dim3(eb, len), 128, 0, st>>>(lag, nf, n, w.tab, E, other, w.M0, w.g0);
        SPYB_LAUNCH_CHECK("plus_causal_kernel"); count_launch();
        zd* gp = nullptr;
        if (fft_axis0(other, lag, w.tw, len, E, plan, &gp, st)) return 1;
        zd* M = w.X;
        SPYB_CUDA(cudaMemsetAsync(w.err_bits, 0, sizeof(unsigned long long), st));
        if (nfl > 0) {
            plus_unpack_kernel<<<dim3(eb, nfl), 128, 0, st>>>(gp, nf, n, w.tab, E, w.g0, M, f_lo);
            SPYB_LAUNCH_CHECK("plus_unpack_kernel"); count_launch();
            // psi <- psi (gplus + S)                         (wilson_sf.py:101)
            if (run_gemm<0, 0>(psi + so, n2, M + so, n2, psi_next + so, n2, nullptr, 0, nullptr, n, nfl, false, st)) return 1;
        }
        // psi0 <- psi0 (gplus_0 + S)                         (wilson_sf.py:102), replicated
        if (run_gemm<0, 0>(psi0, 0, w.M0, 0, psi0_next, 0, nullptr, 0, nullptr, n, 1, false, st)) return 1;
        { zd* t = psi; psi = psi_next; psi_next = t; }
        { zd* t = psi0; psi0 = psi0_next; psi0_next = t; }
        // err = max |CSD - psi psi^H| / |CSD|                 (wilson_sf.py:104-106)
        // S and psi psi^H are Hermitian: the upper tiles carry every value of |S - psi psi^H| / |S|
        if (nfl > 0 && run_gemm<1, 1>(psi + so, n2, psi + so, n2, nullptr, 0, S + so, n2, w.err_bits, n, nfl, true, st)) return 1;
        if (exchange && exchange(exchange_ctx, 1, w.err_bits, 8, 1))
            return fail("wilson_sf: the exchange callback failed while reducing the error");
        unsigned long long bits = 0;
        SPYB_CUDA(cudaMemcpyAsync(&bits, w.err_bits, sizeof(bits), cudaMemcpyDeviceToHost, st));
        SPYB_CUDA(cudaStreamSynchronize(st));
        memcpy(&err, &bits, sizeof(err));
        if (err < rtol) { converged = true; ++it; break; }
    }
    if (agree_failure(nfl > 0 ? check_info(w.info + f_lo, nfl, "wilson_sf (inverse of psi)", st) : 0)) return 1;
    // Sigma = psi0 psi0^T, H = psi psi0^-1                   (wilson_sf.py:114-118)
    if (run_gemm<1, 0>(psi0, 0, psi0, 0, w.tmp0, 0, nullptr, 0, nullptr, n, 1, false, st)) return 1;
    real_part_kernel<<<nb2, 256, 0, st>>>(w.tmp0, Sigma_out, n2);
    SPYB_LAUNCH_CHECK("real_part_kernel"); count_launch();
    if (run_gesv(psi0, 0, w.eye, 0, w.tmp0, 0, w.inv0, 0, n, 1, w.info + nf, st)) return 1;
    if (check_info(w.info + nf, 1, "wilson_sf (inverse of psi0)", st)) return 1;
    if (nfl > 0 && run_gemm<0, 0>(psi + so, n2, w.inv0, 0, static_cast<zd*>(H_out) + so, n2, nullptr, 0, nullptr, n, nfl, false, st))
        return 1;
    *converged_host = converged ? 1 : 0;
    *err_host = err;
    if (iters_host) *iters_host = it;
    return 0;
}

int granger(const void* csd_c128, const void* H, const double* Sigma, int n_freq, int n_chan, float* out,
            cudaStream_t st) {
    const long long total = (long long)n_freq * n_chan * n_chan;
    if (total <= 0) return 0;
    granger_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(static_cast<const zd*>(csd_c128),
                                                                    static_cast<const zd*>(H), Sigma, n_freq, n_chan, out);
    SPYB_LAUNCH_CHECK("granger_kernel");
    count_launch();
    return 0;
}

}  // namespace spyb
