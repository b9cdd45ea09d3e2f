// K2 (SIMT variant) / K3: cross-spectral contraction and coherency normalisation.
//
// K2 replaces the broadcasting outer product + taper mean of
//   syncopy/connectivity/csd.py:98-102 and ST_compRoutines.py:84-116,
// and -- when rows of several trials are handed over at once -- the runtime's trial sum
//   syncopy/shared/computational_routine.py:1022-1032:
//   acc[f][i][j] = beta * acc[f][i][j] + alpha * sum_r X[f][r][si(i)] * conj(X[f][r][sj(j)])
// where r runs over (trial, taper) rows.  Per frequency this is a rank-R Hermitian update;
// this file holds the FP32 CUDA-core version (64x64 tile per block, 4x4 complex per
// thread), csd_tc.cu the tcgen05 tensor-core version used for large aligned problems.
//
// K3 replaces syncopy/connectivity/csd.py:161-170 (coherency) + the output conversion.
#include "common.cuh"
#include "spyb_internal.h"

namespace spyb {

constexpr int CSD_T = 64;      // tile edge
constexpr int CSD_RK = 8;      // rows per stage

struct CsdArgs {
    const float2* X;           // spectra, element (f, r, c) at f*sx_f + r*sx_r + c
    long long sx_f, sx_r;
    int n_rows, n_freq, n_chan;
    const int* idx_i;          // optional channel subsets (spectral_dyadic_product_cF send/rec)
    const int* idx_j;
    int Ci, Cj;
    int hermitian;             // idx_i == idx_j == null: compute upper tiles, mirror the rest
    float2* acc;               // [n_freq][Ci][Cj]
    float alpha, beta;
};

__global__ void __launch_bounds__(256) csd_simt_kernel(const CsdArgs a) {
    __shared__ float2 sA[CSD_RK][CSD_T];
    __shared__ float2 sB[CSD_RK][CSD_T];
    __shared__ float2 sT[CSD_T][CSD_T + 1];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int f = blockIdx.y;

    // decode tile
    int ti, tj;
    const int ntj = (a.Cj + CSD_T - 1) / CSD_T;
    if (a.hermitian) {
        int rem = blockIdx.x;
        ti = 0;
        while (rem >= ntj - ti) { rem -= ntj - ti; ++ti; }
        tj = ti + rem;
    } else {
        ti = blockIdx.x / ntj;
        tj = blockIdx.x % ntj;
    }
    const int i0 = ti * CSD_T, j0 = tj * CSD_T;

    // operand fetch: thread loads 2 elements of A and 2 of B per stage (8 rows x 64 cols each)
    const int lc = tid & 63, lr = tid >> 6;      // lr in 0..3 -> rows lr and lr+4
    int ci = i0 + lc, cj = j0 + lc;
    const bool ci_ok = ci < a.Ci, cj_ok = cj < a.Cj;
    if (a.idx_i && ci_ok) ci = a.idx_i[ci];
    if (a.idx_j && cj_ok) cj = a.idx_j[cj];
    const float2* __restrict__ Xf = a.X + (long long)f * a.sx_f;

    float2 acc[4][4];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[ii][jj] = make_float2(0.f, 0.f);

    const float2 zero = make_float2(0.f, 0.f);
    float2 pa[2], pb[2];
    auto fetch = [&](int r0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = r0 + lr + 4 * h;
            const bool ok = r < a.n_rows;
            pa[h] = (ok && ci_ok) ? __ldg(Xf + (long long)r * a.sx_r + ci) : zero;
            pb[h] = (ok && cj_ok) ? __ldg(Xf + (long long)r * a.sx_r + cj) : zero;
        }
    };

    fetch(0);
    for (int r0 = 0; r0 < a.n_rows; r0 += CSD_RK) {
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            sA[lr + 4 * h][lc] = pa[h];
            sB[lr + 4 * h][lc] = pb[h];
        }
        __syncthreads();
        if (r0 + CSD_RK < a.n_rows) fetch(r0 + CSD_RK);
#pragma unroll
        for (int r = 0; r < CSD_RK; ++r) {
            float2 xi[4], xj[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                xi[q] = sA[r][ty + 16 * q];
                xj[q] = sB[r][tx + 16 * q];
            }
#pragma unroll
            for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    // acc += xi * conj(xj)
                    acc[ii][jj].x = fmaf(xi[ii].x, xj[jj].x, fmaf(xi[ii].y, xj[jj].y, acc[ii][jj].x));
                    acc[ii][jj].y = fmaf(xi[ii].y, xj[jj].x, fmaf(-xi[ii].x, xj[jj].y, acc[ii][jj].y));
                }
        }
    }

    // ---- epilogue ----
    // Hermitian mode: every value above (or on) the diagonal is computed once and written twice
    // (as is, and conjugated into the mirrored position), so the result is exactly Hermitian.
    float2* __restrict__ out = a.acc + (long long)f * a.Ci * a.Cj;
    const bool diag_tile = a.hermitian && ti == tj;
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
        const int i = i0 + ty + 16 * ii;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = j0 + tx + 16 * jj;
            float2 val = make_float2(acc[ii][jj].x * a.alpha, acc[ii][jj].y * a.alpha);
            if (diag_tile && i == j) val.y = 0.f;           // auto-spectra are exactly real
            if (a.hermitian) sT[ty + 16 * ii][tx + 16 * jj] = val;
            if (i < a.Ci && j < a.Cj && (!diag_tile || i <= j)) {
                float2* o = out + (long long)i * a.Cj + j;
                if (a.beta != 0.f) { const float2 old = *o; val.x += a.beta * old.x; val.y += a.beta * old.y; }
                *o = val;
            }
        }
    }
    if (a.hermitian) {
        __syncthreads();
        // mirrored block: out[j][i] = conj(val[i][j]); lanes run along i for coalescing
        for (int e = tid; e < CSD_T * CSD_T; e += 256) {
            const int jl = e >> 6, il = e & 63;
            const int i = i0 + il, j = j0 + jl;
            if (i < a.Ci && j < a.Cj && (!diag_tile || il < jl)) {
                float2 val = sT[il][jl];
                val.y = -val.y;
                float2* o = out + (long long)j * a.Cj + i;
                if (a.beta != 0.f) { const float2 old = *o; val.x += a.beta * old.x; val.y += a.beta * old.y; }
                *o = val;
            }
        }
    }
}

int csd_accumulate_simt(const CsdDesc& d, cudaStream_t stream) {
    if (d.n_freq <= 0 || d.n_chan <= 0) return 0;
    CsdArgs a;
    a.X = reinterpret_cast<const float2*>(d.spectra);
    a.sx_f = d.sx_f; a.sx_r = d.sx_r;
    a.n_rows = d.n_rows; a.n_freq = d.n_freq; a.n_chan = d.n_chan;
    a.idx_i = d.idx_i; a.idx_j = d.idx_j;
    a.Ci = d.idx_i ? d.n_i : d.n_chan;
    a.Cj = d.idx_j ? d.n_j : d.n_chan;
    a.hermitian = (!d.idx_i && !d.idx_j) ? 1 : 0;
    a.acc = reinterpret_cast<float2*>(d.acc);
    a.alpha = d.alpha; a.beta = d.beta;
    const int nti = (a.Ci + CSD_T - 1) / CSD_T, ntj = (a.Cj + CSD_T - 1) / CSD_T;
    const int tiles = a.hermitian ? nti * (nti + 1) / 2 : nti * ntj;
    if (d.n_freq > 65535) return fail("csd: more than 65535 frequencies per launch (%d)", d.n_freq);
    dim3 grid(tiles, d.n_freq);
    csd_simt_kernel<<<grid, 256, 0, stream>>>(a);
    SPYB_LAUNCH_CHECK("csd_simt_kernel");
    count_launch();
    return 0;
}

// ---------------------------------------------------------------------------------------
// K3: coherency  C_ij / sqrt(C_ii C_jj)  (+ scale) and output conversion
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float2 csqrt_principal(float2 z) {
    const float m = hypotf(z.x, z.y);
    if (m == 0.f) return make_float2(0.f, z.y);
    float re = sqrtf(0.5f * (m + fabsf(z.x)));
    float im = 0.5f * z.y / re;
    if (z.x >= 0.f) return make_float2(re, im);
    // z.x < 0: swap roles, keep the sign of the imaginary part
    return make_float2(fabsf(im), copysignf(re, z.y));
}

// One block = NORM_ROWS rows of one matrix.  The diagonal is staged once per block: when it is real and
// non-negative (every CSD this library produces; auto-spectra are exactly real) 1/sqrt(pre*C_jj) is kept in
// shared memory and an element costs two multiplies; otherwise the general complex square root of the
// reference (np.sqrt of the complex product, csd.py:161-170) is taken.  Rows are read / written as whole
// 128-byte lines (two complex per thread and step).
constexpr int NORM_ROWS = 8;

template <int KIND>
__device__ __forceinline__ void norm_store(void* __restrict__ out, long long o, float2 c0, float2 c1) {
    if (KIND == OUT_FOURIER) {
        reinterpret_cast<float4*>(out)[o >> 1] = make_float4(c0.x, c0.y, c1.x, c1.y);
    } else {
        float r0, r1;
        if (KIND == OUT_ABS) { r0 = sqrtf(c0.x * c0.x + c0.y * c0.y); r1 = sqrtf(c1.x * c1.x + c1.y * c1.y); }
        else { r0 = convert_real(c0, KIND); r1 = convert_real(c1, KIND); }
        reinterpret_cast<float2*>(out)[o >> 1] = make_float2(r0, r1);
    }
}

template <int KIND>
__global__ void __launch_bounds__(256) csd_normalize_kernel(const float2* __restrict__ csd, int n_chan,
                                                            long long n_mat, float pre_scale,
                                                            void* __restrict__ out) {
    extern __shared__ float2 s_diag[];           // [n_chan] scaled diagonal, then [n_chan] floats 1/sqrt
    float* s_rs = reinterpret_cast<float*>(s_diag + n_chan);
    __shared__ int s_general;
    const int row_blocks = (n_chan + NORM_ROWS - 1) / NORM_ROWS;
    const long long mat = blockIdx.x / row_blocks;
    const int i0 = (int)(blockIdx.x % row_blocks) * NORM_ROWS;
    const float2* __restrict__ M = csd + mat * n_chan * n_chan;
    if (threadIdx.x == 0) s_general = 0;
    __syncthreads();
    for (int j = threadIdx.x; j < n_chan; j += blockDim.x) {
        float2 d = M[(long long)j * n_chan + j];
        d.x *= pre_scale; d.y *= pre_scale;
        s_diag[j] = d;
        s_rs[j] = rsqrtf(d.x);
        if (d.y != 0.f || !(d.x > 0.f)) s_general = 1;
    }
    __syncthreads();
    const bool general = s_general != 0;
    const bool vec = (n_chan % 2) == 0;
    for (int r = 0; r < NORM_ROWS; ++r) {
        const int i = i0 + r;
        if (i >= n_chan) break;
        const float2 dii = s_diag[i];
        const float rsi = s_rs[i] * pre_scale;
        const long long row = (mat * n_chan + i) * n_chan;
        for (int j = 2 * threadIdx.x; j < n_chan; j += 2 * blockDim.x) {
            float2 c[2];
            const bool two = j + 1 < n_chan;
            if (vec) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(csd + row + j));
                c[0] = make_float2(v.x, v.y); c[1] = make_float2(v.z, v.w);
            } else {
                c[0] = __ldg(csd + row + j);
                c[1] = two ? __ldg(csd + row + j + 1) : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int jj = j + e < n_chan ? j + e : j;
                if (!general) {
                    const float w = rsi * s_rs[jj];
                    c[e].x *= w; c[e].y *= w;
                } else {
                    const float2 num = make_float2(c[e].x * pre_scale, c[e].y * pre_scale);
                    const float2 den = csqrt_principal(cmul(dii, s_diag[jj]));
                    const float d2 = den.x * den.x + den.y * den.y;
                    c[e] = make_float2((num.x * den.x + num.y * den.y) / d2, (num.y * den.x - num.x * den.y) / d2);
                }
            }
            if (vec) {
                norm_store<KIND>(out, row + j, c[0], c[1]);
            } else {
                for (int e = 0; e < (two ? 2 : 1); ++e) {
                    if (KIND == OUT_FOURIER) reinterpret_cast<float2*>(out)[row + j + e] = c[e];
                    else reinterpret_cast<float*>(out)[row + j + e] =
                        KIND == OUT_ABS ? sqrtf(c[e].x * c[e].x + c[e].y * c[e].y) : convert_real(c[e], KIND);
                }
            }
        }
    }
}

template <int KIND>
static int launch_normalize(const void* csd, long long n_mat, int n_chan, float pre_scale, void* out,
                            cudaStream_t stream) {
    const long long blocks = n_mat * ((n_chan + NORM_ROWS - 1) / NORM_ROWS);
    if (blocks > 2147483647LL) return fail("csd_normalize: too many blocks (%lld)", blocks);
    const int threads = n_chan >= 512 ? 256 : (n_chan >= 256 ? 128 : 64);
    const size_t smem = (size_t)n_chan * (sizeof(float2) + sizeof(float));
    if (smem > 48 * 1024) return fail("csd_normalize: more than 4096 channels are not supported");
    csd_normalize_kernel<KIND><<<(unsigned)blocks, threads, smem, stream>>>(
        reinterpret_cast<const float2*>(csd), n_chan, n_mat, pre_scale, out);
    SPYB_LAUNCH_CHECK("csd_normalize_kernel");
    count_launch();
    return 0;
}

int csd_normalize(const void* csd, long long n_mat, int n_chan, float pre_scale, int out_kind, void* out,
                  cudaStream_t stream) {
    if (n_mat <= 0 || n_chan <= 0) return 0;
    switch (out_kind) {
        case OUT_POW:     return launch_normalize<OUT_POW>(csd, n_mat, n_chan, pre_scale, out, stream);
        case OUT_ABS:     return launch_normalize<OUT_ABS>(csd, n_mat, n_chan, pre_scale, out, stream);
        case OUT_FOURIER: return launch_normalize<OUT_FOURIER>(csd, n_mat, n_chan, pre_scale, out, stream);
        case OUT_REAL:    return launch_normalize<OUT_REAL>(csd, n_mat, n_chan, pre_scale, out, stream);
        case OUT_IMAG:    return launch_normalize<OUT_IMAG>(csd, n_mat, n_chan, pre_scale, out, stream);
        case OUT_ANGLE:   return launch_normalize<OUT_ANGLE>(csd, n_mat, n_chan, pre_scale, out, stream);
        case OUT_ABSREAL: return launch_normalize<OUT_ABSREAL>(csd, n_mat, n_chan, pre_scale, out, stream);
        case OUT_ABSIMAG: return launch_normalize<OUT_ABSIMAG>(csd, n_mat, n_chan, pre_scale, out, stream);
        default:          return fail("bad out_kind %d", out_kind);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Coherency straight from the "tile slots" the tcgen05 kernel fills (csd_tc.cu, store_mode 2):
//   slots [n_src][n_freq][n_tiles][128][128] complex64, one slot set per source rank, upper tiles only.
// Sums the source ranks (fixed order: deterministic), applies pre_scale (1 / nTrials), normalises with the
// diagonal and writes the full Hermitian-consistent [n_freq][C][C] result: elements on / above the diagonal from
// the tile, the mirrored ones as their conjugates (csd.py:118-172 on the trial average,
// computational_routine.py:1022-1032 for the sum over ranks).
// ---------------------------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ void conv_store(void* out, long long idx, float2 c) {
    if (KIND == OUT_FOURIER) reinterpret_cast<float2*>(out)[idx] = c;
    else reinterpret_cast<float*>(out)[idx] = KIND == OUT_ABS ? sqrtf(c.x * c.x + c.y * c.y) : convert_real(c, KIND);
}

template <int KIND>
__global__ void __launch_bounds__(256) csd_normalize_tiles_kernel(const float2* __restrict__ slots, int n_src,
                                                                  long long src_stride, int n_tiles, int C,
                                                                  float pre_scale, void* __restrict__ out) {
    // block = 32 rows x 128 columns of one tile, walked as 32 x 32 chunks; the mirrored half of every chunk goes
    // through a shared-memory transpose so that both halves are stored as 128-byte row pieces
    __shared__ float2 s_di[32], s_dj[128];
    __shared__ float s_ri[32], s_rj[128];
    __shared__ float2 s_t[32][33];
    __shared__ int s_general;
    const int t = blockIdx.y, f = blockIdx.z;
    const int nb = (C + 127) / 128;                              // the last block may be zero padding (C % 32 == 0)
    int ti, tj;
    tri_decode(nb, t, ti, tj);
    const int bi = blockIdx.x;
    if (ti * 128 + bi * 32 >= C) return;                         // padding rows (block-uniform)
    const bool diag_tile = ti == tj;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const float2* __restrict__ fbase = slots + (long long)f * n_tiles * (128 * 128);
    const float2* __restrict__ tile = fbase + t * (128 * 128);
    if (tid == 0) s_general = 0;
    __syncthreads();
    if (tid < 160) {
        const bool col = tid >= 32;
        const int l = col ? tid - 32 : bi * 32 + tid;
        const int tt = tri_diag(nb, col ? tj : ti);
        const float2* __restrict__ dt = fbase + tt * (128 * 128) + l * 129;
        float2 d = make_float2(0.f, 0.f);
        for (int sidx = 0; sidx < n_src; ++sidx) { const float2 w = __ldg(dt + sidx * src_stride); d.x += w.x; d.y += w.y; }
        d.x *= pre_scale; d.y *= pre_scale;
        const float rs = rsqrtf(d.x);
        if (col) { s_dj[tid - 32] = d; s_rj[tid - 32] = rs; } else { s_di[tid] = d; s_ri[tid] = rs; }
        const int gl = (col ? tj : ti) * 128 + l;                // padding channels have a zero diagonal: not a reason
        if (gl < C && (d.y != 0.f || !(d.x > 0.f))) s_general = 1;   // for the general (complex-sqrt) path
    }
    __syncthreads();
    const bool general = s_general != 0;
    const int I0 = ti * 128 + bi * 32;
    float* __restrict__ outf = reinterpret_cast<float*>(out) + (long long)f * C * C * (KIND == OUT_FOURIER ? 2 : 1);
    float2* __restrict__ outc = reinterpret_cast<float2*>(outf);

    for (int bj = diag_tile ? bi : 0; bj < 4; ++bj) {            // chunks left of the diagonal are never read
        const bool diag_blk = diag_tile && bi == bj;
        const int J0 = tj * 128 + bj * 32;
        if (J0 >= C) break;                                      // padding columns
        float2 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = make_float2(0.f, 0.f);
        const float2* __restrict__ src = tile + (bi * 32 + ty) * 128 + bj * 32 + tx;
        for (int sidx = 0; sidx < n_src; ++sidx) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 w = __ldg(src + k * (8 * 128));
                v[k].x += w.x; v[k].y += w.y;
            }
            src += src_stride;
        }
        const float rsj = s_rj[bj * 32 + tx] * pre_scale;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = ty + 8 * k;
            float2 c;
            if (!general) {
                const float w = s_ri[r] * rsj;
                c = make_float2(v[k].x * w, v[k].y * w);
                if (diag_blk && tx == r) c.y = 0.f;
            } else {
                const float2 x = make_float2(v[k].x * pre_scale, v[k].y * pre_scale);
                const float2 den = csqrt_principal(cmul(s_di[r], s_dj[bj * 32 + tx]));
                const float d2 = den.x * den.x + den.y * den.y;
                c = make_float2((x.x * den.x + x.y * den.y) / d2, (x.y * den.x - x.x * den.y) / d2);
            }
            // convert once; the mirrored element is the converted conjugate (bit-identical magnitudes)
            const int idx = (I0 + r) * C + J0 + tx;
            if (KIND == OUT_FOURIER) {
                if (!diag_blk || tx >= r) outc[idx] = c;
                s_t[r][tx] = make_float2(c.x, -c.y);
            } else {
                float val;
                if (KIND == OUT_ABS) { const float q = c.x * c.x + c.y * c.y; val = q > 0.f ? q * rsqrtf(q) : 0.f; }
                else val = convert_real(c, KIND);
                if (!diag_blk || tx >= r) outf[idx] = val;
                s_t[r][tx].x = (KIND == OUT_IMAG || KIND == OUT_ANGLE) ? -val : val;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int jr = ty + 8 * k;                           // source column = output row
            if (!diag_blk || jr > tx) {                          // source (i = tx, j = jr) strictly above the diagonal
                const int idx = (J0 + jr) * C + I0 + tx;
                if (KIND == OUT_FOURIER) outc[idx] = s_t[tx][jr];
                else outf[idx] = s_t[tx][jr].x;
            }
        }
        __syncthreads();
    }
}

template <int KIND>
static int launch_normalize_tiles(const void* slots, int n_src, int n_freq, int n_chan, float pre_scale, void* out,
                                  cudaStream_t stream) {
    const int n_tiles = tri_tiles((n_chan + 127) / 128);
    const long long src_stride = (long long)n_freq * n_tiles * 128 * 128;
    dim3 grid(4, n_tiles, n_freq);
    csd_normalize_tiles_kernel<KIND><<<grid, 256, 0, stream>>>(reinterpret_cast<const float2*>(slots), n_src, src_stride,
                                                               n_tiles, n_chan, pre_scale, out);
    SPYB_LAUNCH_CHECK("csd_normalize_tiles_kernel");
    count_launch();
    return 0;
}

int csd_normalize_tiles(const void* slots, int n_src, int n_freq, int n_chan, float pre_scale, int out_kind,
                        void* out, cudaStream_t stream) {
    if (n_freq <= 0) return 0;
    if (n_chan < 64 || n_chan > 512 || n_chan % 32) return fail("tile slots exist for 64..512 channels in multiples of 32 (got %d)", n_chan);
    if (n_src < 1) return fail("csd_normalize_tiles: need at least one source slot");
    if (n_freq > 65535) return fail("csd_normalize_tiles: more than 65535 frequencies per call are not supported");
    switch (out_kind) {
        case OUT_POW:     return launch_normalize_tiles<OUT_POW>(slots, n_src, n_freq, n_chan, pre_scale, out, stream);
        case OUT_ABS:     return launch_normalize_tiles<OUT_ABS>(slots, n_src, n_freq, n_chan, pre_scale, out, stream);
        case OUT_FOURIER: return launch_normalize_tiles<OUT_FOURIER>(slots, n_src, n_freq, n_chan, pre_scale, out, stream);
        case OUT_REAL:    return launch_normalize_tiles<OUT_REAL>(slots, n_src, n_freq, n_chan, pre_scale, out, stream);
        case OUT_IMAG:    return launch_normalize_tiles<OUT_IMAG>(slots, n_src, n_freq, n_chan, pre_scale, out, stream);
        case OUT_ANGLE:   return launch_normalize_tiles<OUT_ANGLE>(slots, n_src, n_freq, n_chan, pre_scale, out, stream);
        case OUT_ABSREAL: return launch_normalize_tiles<OUT_ABSREAL>(slots, n_src, n_freq, n_chan, pre_scale, out, stream);
        case OUT_ABSIMAG: return launch_normalize_tiles<OUT_ABSIMAG>(slots, n_src, n_freq, n_chan, pre_scale, out, stream);
        default:          return fail("bad out_kind %d", out_kind);
    }
}

// in-place scale of a float buffer (trial mean: computational_routine.py:1030-1032)
// Lower triangle <- conjugate of the upper one, imaginary part of the diagonal <- 0, in place on [n_freq][C][C]
// complex64.  A sum of exactly Hermitian partial matrices over ranks is not exactly Hermitian any more when the
// collective adds the partials of (i, j) and of (j, i) in different orders (ring all-reduce: the order depends on
// the chunk an element falls into); the Wilson iteration's element-wise error then stalls at that asymmetry.
__global__ void __launch_bounds__(256) mirror_upper_kernel(float2* __restrict__ csd, int n) {
    float2* __restrict__ m = csd + (long long)blockIdx.z * n * n;
    const int j = blockIdx.x * 32 + (threadIdx.x & 31);
    const int i = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (i >= n || j >= n || j > i) return;
    if (j == i) { m[(long long)i * n + i].y = 0.f; return; }
    const float2 u = m[(long long)j * n + i];
    m[(long long)i * n + j] = make_float2(u.x, -u.y);
}

int csd_mirror_upper(void* csd, int n_freq, int n_chan, cudaStream_t stream) {
    if (n_freq <= 0 || n_chan <= 0) return 0;
    if (n_freq > 65535) return fail("csd_mirror_upper: too many frequencies per launch (%d)", n_freq);
    dim3 grid((n_chan + 31) / 32, (n_chan + 7) / 8, n_freq);
    mirror_upper_kernel<<<grid, 256, 0, stream>>>(static_cast<float2*>(csd), n_chan);
    SPYB_LAUNCH_CHECK("mirror_upper_kernel");
    count_launch();
    return 0;
}

__global__ void scale_kernel(float* __restrict__ x, long long n, float s) {
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = i0; i < n; i += stride) x[i] *= s;
}

int scale_inplace(float* x, long long n, float s, cudaStream_t stream) {
    if (n <= 0) return 0;
    long long blocks = (n + 1023) / 1024;
    if (blocks > 148 * 16) blocks = 148 * 16;
    scale_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, n, s);
    SPYB_LAUNCH_CHECK("scale_kernel");
    count_launch();
    return 0;
}

// acc[e] = beta * acc[e] + alpha * sum_b src[b][e]: the runtime's `target[()] += res` over a batch of trials and the
// final `/= numTrials` (computational_routine.py:1025,1030-1032).  Trials are added in index order per element.
__global__ void sum_trials_kernel(const float4* __restrict__ src, int n_trials, long long stride4, long long n4,
                                  float alpha, float beta, float4* __restrict__ acc) {
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = i0; i < n4; i += step) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int b = 0; b < n_trials; ++b) {
            const float4 v = __ldg(src + (long long)b * stride4 + i);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        float4 o = make_float4(alpha * s.x, alpha * s.y, alpha * s.z, alpha * s.w);
        if (beta != 0.f) {
            const float4 a = acc[i];
            o.x += beta * a.x; o.y += beta * a.y; o.z += beta * a.z; o.w += beta * a.w;
        }
        acc[i] = o;
    }
}

int sum_trials(const float* src, int n_trials, long long trial_stride, long long n_elems, float alpha, float beta,
               float* acc, cudaStream_t stream) {
    if (n_elems <= 0) return 0;
    if (n_elems % 4 || trial_stride % 4 || reinterpret_cast<uintptr_t>(src) % 16 || reinterpret_cast<uintptr_t>(acc) % 16)
        return fail("sum_trials: element count, trial stride and pointers must be multiples of 4 floats / 16 bytes");
    const long long n4 = n_elems / 4;
    long long blocks = (n4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    sum_trials_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(src), n_trials, trial_stride / 4, n4,
                                                            alpha, beta, reinterpret_cast<float4*>(acc));
    SPYB_LAUNCH_CHECK("sum_trials_kernel");
    count_launch();
    return 0;
}

}  // namespace spyb
