// K5 / K6: wavelet and superlet transforms as FFT convolutions, plus the whole-trial detrend they need.
//
// Replaces the NumPy/SciPy bodies of
//   syncopy/specest/wavelets/transform.py:88-108   cwt_time: fftconvolve(data, psi_s, 'same') per scale
//   syncopy/specest/superlet.py:108-198, 321-365   cwtSL per (cycle count, scale) + complex geometric means
//   syncopy/specest/compRoutines.py:582-595, 751-762 (detrend, output conversion)
//
// 'same' convolution with a sampled kernel psi (length M, centre c0 = (M-1)/2):
//   out[n] = sum_m x[m] psi[n - m + c0],  0 <= n < N
// equals the circular convolution of length L >= N + max(c0, M-1-c0) of the zero-padded trial with
// h[d mod L] = psi[d + c0].  The host samples psi in float64 exactly as the reference does and hands over
// T[s][j][k] = FFT_L(h)[k] / L per (scale s, factor j); the forward spectra X[c][k] (k <= L/2) of the
// trials come from the mtmfft kernel.  Here one block inverse-transforms X_c * T_sj for P channels:
//   y = conj(FFT_L(conj(X T)))   (same Stockham block FFT as everywhere else)
// and folds the factors of a scale as the reference's superlets do:
//   z[n] = prod_j y_j[n] ^ a_sj     (principal branch complex power, np.power on complex64)
// A plain wavelet transform is the one-factor case with exponent 1.
#include "common.cuh"
#include "fft_core.cuh"
#include "plan.cuh"
#include "spyb_internal.h"

namespace spyb {

// ---------------------------------------------------------------------------------------
// whole-trial detrend (scipy.signal.detrend type 'constant' / 'linear' along time), float32
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) detrend_kernel(const float* __restrict__ x, long long trial_stride,
                                                      int n_samples, int n_chan, int polyremoval,
                                                      float* __restrict__ out, long long out_trial_stride) {
    // block = 32 channels x 8 time lanes of one trial
    __shared__ float red[2][8][33];
    const int cl = threadIdx.x & 31, tl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const float* __restrict__ xt = x + (long long)blockIdx.y * trial_stride;
    float* __restrict__ ot = out + (long long)blockIdx.y * out_trial_stride;
    const float tmid = 0.5f * (float)(n_samples - 1);
    float s0 = 0.f, s1 = 0.f;
    if (c < n_chan) {
        for (int n = tl; n < n_samples; n += 8) {
            const float v = xt[(long long)n * n_chan + c];
            s0 += v;
            s1 += ((float)n - tmid) * v;
        }
    }
    red[0][tl][cl] = s0; red[1][tl][cl] = s1;
    __syncthreads();
    float mean = 0.f, slope = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) { mean += red[0][t][cl]; slope += red[1][t][cl]; }
    mean /= (float)n_samples;
    if (polyremoval == 1 && n_samples > 1)
        slope /= (float)((double)n_samples * ((double)n_samples * n_samples - 1.0) / 12.0);
    else
        slope = 0.f;
    if (polyremoval < 0) mean = 0.f;
    if (c < n_chan) {
        for (int n = tl; n < n_samples; n += 8) {
            const long long o = (long long)n * n_chan + c;
            ot[o] = xt[o] - (mean + slope * ((float)n - tmid));
        }
    }
}

int detrend(const float* x, int n_trials, long long trial_stride, int n_samples, int n_chan, int polyremoval,
            float* out, long long out_trial_stride, cudaStream_t stream) {
    if (n_trials <= 0 || n_samples <= 0 || n_chan <= 0) return 0;
    if (n_trials > 65535) return fail("detrend: too many trials per launch (%d)", n_trials);
    dim3 grid((n_chan + 31) / 32, n_trials);
    detrend_kernel<<<grid, 256, 0, stream>>>(x, trial_stride, n_samples, n_chan, polyremoval, out, out_trial_stride);
    SPYB_LAUNCH_CHECK("detrend_kernel");
    count_launch();
    return 0;
}

// ---------------------------------------------------------------------------------------
// K5 / K6
// ---------------------------------------------------------------------------------------
struct CwtArgs {
    const float2* xspec;       // [trial][L/2+1][chan] one-sided spectra of the zero-padded trials
    int n_trials, n_chan, n_dft;
    const float2* kern;        // [scale][max_fac][L]  FFT_L(h) / L
    const float* expo;         // [scale][max_fac]
    const int* n_fac;          // [scale]
    int n_scales, max_fac;
    int n_time;                // rows written: n = 0 .. n_time-1
    int out_kind;
    void* out;                 // [trial][n_time][scale][chan]
    int transposed;            // 1: xspec is [trial][chan][L/2+1] and out is [trial][scale][chan][n_time] -- every
                               // global access of the kernel is then contiguous across the threads of a warp
    const float2* tw;
};

// z ^ a, principal branch (np.power(complex64, real))
__device__ __forceinline__ float2 cpow_real(float2 z, float a) {
    const float r = hypotf(z.x, z.y);
    if (r == 0.f) return make_float2(0.f, 0.f);
    const float mag = powf(r, a);
    float sn, cs;
    sincosf(a * atan2f(z.y, z.x), &sn, &cs);
    return make_float2(mag * cs, mag * sn);
}

template <int LOG2N, int P, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) cwt_kernel(const CwtArgs a) {
    constexpr int N = 1 << LOG2N;
    constexpr int NT = N / 16;
    constexpr int SPAD = fft_padded_len(N);
    extern __shared__ float2 smem[];
    float2* s = smem;                       // exchange buffer [SPAD][P]
    float2* zacc = smem + (size_t)SPAD * P; // running product [n_time][P] (only when a scale has > 1 factor)

    const int tid = threadIdx.x;
    const int p = tid % P;
    const int j = tid / P;
    const int sc = blockIdx.y, trial = blockIdx.z;
    const int c = blockIdx.x * P + p;
    const bool c_ok = c < a.n_chan;
    const int nf = a.n_fac[sc];
    const int half = N / 2;
    const float2* __restrict__ X = a.xspec + (long long)trial * (half + 1) * a.n_chan;
    const long long xs = a.transposed ? 1 : a.n_chan;                               // stride between bins
    const long long xc = a.transposed ? (long long)c * (half + 1) : (long long)c;   // offset of the channel

    // superlets with a magnitude output only need |z|: the geometric mean runs on log-magnitudes
    const bool mag_only = nf > 1 && (a.out_kind == OUT_POW || a.out_kind == OUT_ABS);
    float2 v[16];
    for (int fj = 0; fj < nf; ++fj) {
        const float2* __restrict__ T = a.kern + ((long long)sc * a.max_fac + fj) * N;
        const float ex = a.expo[sc * a.max_fac + fj];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int k = j + NT * e;
            float2 xv = make_float2(0.f, 0.f);
            if (c_ok) {
                if (k <= half) xv = __ldg(X + (long long)k * xs + xc);
                else { xv = __ldg(X + (long long)(N - k) * xs + xc); xv.y = -xv.y; }
            }
            v[e] = cconj(cmul(xv, __ldg(T + k)));
        }
        block_fft<LOG2N, P>(v, s, a.tw, j, p);          // natural order result in s (ends with a barrier)
        if (nf > 1) {
            if (mag_only) {
                // |prod_j y_j^a_j| = exp2(sum_j a_j log2|y_j|): no phases, no complex powers (pow / abs outputs)
                for (int n = j; n < a.n_time; n += NT) {
                    const float2 y = s[fft_pad(n) * P + p];
                    const float l = 0.5f * ex * __log2f(y.x * y.x + y.y * y.y);
                    zacc[n * P + p].x = fj == 0 ? l : zacc[n * P + p].x + l;
                }
            } else {
                for (int n = j; n < a.n_time; n += NT) {
                    float2 y = cconj(s[fft_pad(n) * P + p]);
                    if (ex != 1.f) y = cpow_real(y, ex);
                    zacc[n * P + p] = fj == 0 ? y : cmul(zacc[n * P + p], y);
                }
            }
            __syncthreads();                             // exchange buffer is rewritten by the next factor
        }
    }

    // ---- output: element (n, scale, chan) of this trial ----
    if (!c_ok) return;
    const long long row = a.transposed ? 1 : (long long)a.n_scales * a.n_chan;
    const long long base = a.transposed
        ? (((long long)trial * a.n_scales + sc) * a.n_chan + c) * a.n_time
        : (long long)trial * a.n_time * a.n_scales * a.n_chan + (long long)sc * a.n_chan + c;
    for (int n = j; n < a.n_time; n += NT) {
        float2 z;
        if (mag_only) {
            const float mag = exp2f(zacc[n * P + p].x);
            reinterpret_cast<float*>(a.out)[base + (long long)n * row] = a.out_kind == OUT_POW ? mag * mag : mag;
            continue;
        }
        if (nf > 1) z = zacc[n * P + p];
        else {
            z = cconj(s[fft_pad(n) * P + p]);
            const float ex = a.expo[sc * a.max_fac];
            if (ex != 1.f) z = cpow_real(z, ex);
        }
        const long long o = base + (long long)n * row;
        if (a.out_kind == OUT_FOURIER) reinterpret_cast<float2*>(a.out)[o] = z;
        else reinterpret_cast<float*>(a.out)[o] = convert_real(z, a.out_kind);
    }
}

template <int LOG2N, int P>
static int launch_cwt(const CwtArgs& a, bool need_acc, cudaStream_t stream) {
    constexpr int N = 1 << LOG2N, NT = N / 16, THREADS = NT * P;
    auto kern = cwt_kernel<LOG2N, P, THREADS>;
    const size_t smem = (size_t)fft_padded_len(N) * P * sizeof(float2) +
                        (need_acc ? (size_t)a.n_time * P * sizeof(float2) : 0);
    if (smem > 227 * 1024) return fail("cwt: transform does not fit shared memory (%zu bytes)", smem);
    // per device / context attribute: set on every launch (several engines may live in one process)
    SPYB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (a.n_scales > 65535 || a.n_trials > 65535) return fail("cwt: too many scales / trials per launch");
    dim3 grid((a.n_chan + P - 1) / P, a.n_scales, a.n_trials);
    kern<<<grid, THREADS, smem, stream>>>(a);
    SPYB_LAUNCH_CHECK("cwt_kernel");
    count_launch();
    return 0;
}

int cwt_factors(const CwtDesc& d, cudaStream_t stream) {
    if (d.n_trials <= 0 || d.n_chan <= 0 || d.n_scales <= 0 || d.n_time <= 0) return 0;
    if (d.n_time > d.n_dft) return fail("cwt: n_time (%d) exceeds the transform length (%d)", d.n_time, d.n_dft);
    if (d.max_fac < 1) return fail("cwt: need at least one factor per scale");
    if (d.n_dft > 16384) return cwt_factors_long(d, stream);          // beyond the shared-memory kernel
    if (d.n_dft < 16 || (d.n_dft & (d.n_dft - 1))) return fail("cwt: transform length must be a power of two >= 16");
    if (d.n_time > d.n_dft) return fail("cwt: n_time (%d) exceeds the transform length (%d)", d.n_time, d.n_dft);
    if (d.max_fac < 1) return fail("cwt: need at least one factor per scale");
    const FftPlan* pl = get_fft_plan(d.n_dft);
    if (!pl) return 1;
    CwtArgs a;
    a.xspec = reinterpret_cast<const float2*>(d.xspec);
    a.n_trials = d.n_trials; a.n_chan = d.n_chan; a.n_dft = d.n_dft;
    a.kern = reinterpret_cast<const float2*>(d.kern);
    a.expo = d.expo; a.n_fac = d.n_fac;
    a.n_scales = d.n_scales; a.max_fac = d.max_fac;
    a.n_time = d.n_time; a.out_kind = d.out_kind; a.out = d.out;
    a.transposed = d.transposed;
    a.tw = pl->tw;
    const bool acc = d.max_fac > 1;
    switch (pl->log2n) {
        case 4:  return launch_cwt<4, 4>(a, acc, stream);
        case 5:  return launch_cwt<5, 4>(a, acc, stream);
        case 6:  return launch_cwt<6, 4>(a, acc, stream);
        case 7:  return launch_cwt<7, 4>(a, acc, stream);
        case 8:  return launch_cwt<8, 4>(a, acc, stream);
        case 9:  return launch_cwt<9, 4>(a, acc, stream);
        case 10: return launch_cwt<10, 4>(a, acc, stream);
        case 11: return launch_cwt<11, 4>(a, acc, stream);
        case 12: return launch_cwt<12, 2>(a, acc, stream);
        case 13: return launch_cwt<13, 2>(a, acc, stream);
        case 14: return launch_cwt<14, 1>(a, acc, stream);
        default: return fail("cwt: transform length 2^%d not supported (max 2^14)", pl->log2n);
    }
}

// ---------------------------------------------------------------------------------------
// batched 2-D transpose of 4- or 8-byte elements: in [b][rows][cols] -> out [b][cols][rows] (32 x 32 tiles through
// padded shared memory, both sides coalesced); brings spectra / results into and out of the layouts above
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ in, T* __restrict__ out, int rows, int cols) {
    __shared__ T tile[32][33];
    const long long b = blockIdx.z;
    const T* __restrict__ ib = in + b * rows * cols;
    T* __restrict__ ob = out + b * rows * cols;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + 8 * i, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + 8 * i][tx] = ib[(long long)r * cols + c];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, r = r0 + tx;
        if (r < rows && c < cols) ob[(long long)c * rows + r] = tile[tx][ty + 8 * i];
    }
}

int transpose2d(const void* in, void* out, int batch, int rows, int cols, int elem_bytes, cudaStream_t stream) {
    if (batch <= 0 || rows <= 0 || cols <= 0) return 0;
    if (elem_bytes != 4 && elem_bytes != 8) return fail("transpose: element size must be 4 or 8 bytes (got %d)", elem_bytes);
    const unsigned gy = (unsigned)((rows + 31) / 32);
    if (batch > 65535 || gy > 65535) return fail("transpose: too many batches / rows per launch");
    dim3 grid((cols + 31) / 32, gy, batch);
    if (elem_bytes == 4)
        transpose_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(in), static_cast<float*>(out), rows, cols);
    else
        transpose_kernel<float2><<<grid, 256, 0, stream>>>(static_cast<const float2*>(in), static_cast<float2*>(out), rows, cols);
    SPYB_LAUNCH_CHECK("transpose_kernel");
    count_launch();
    return 0;
}

// Transpose with a placement: in [nb1 * nb2][rows][cols] -> out element (b1, b2, col, row) at
// b1 * stride_b1 + b2 * stride_b2 + col * ld_out + row, columns of segment b2 kept while b2 * cols + col < col_limit.
// Lands the time-contiguous rows of a (segmented) wavelet launch in its slice [time][scale range][chan] of the result.
template <typename T>
__global__ void __launch_bounds__(256) transpose_place_kernel(const T* __restrict__ in, T* __restrict__ out, int nb2, int rows,
                                                              int cols, long long stride_b1, long long stride_b2,
                                                              long long ld_out, int col_limit) {
    __shared__ T tile[32][33];
    const int b = blockIdx.z, b1 = b / nb2, b2 = b - b1 * nb2;
    const T* __restrict__ ib = in + (long long)b * rows * cols;
    T* __restrict__ ob = out + b1 * stride_b1 + b2 * stride_b2;
    const int keep = min(cols, col_limit - b2 * cols);
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    if (c0 >= keep) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + 8 * i, c = c0 + tx;
        if (r < rows && c < keep) tile[ty + 8 * i][tx] = ib[(long long)r * cols + c];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, r = r0 + tx;
        if (r < rows && c < keep) ob[(long long)c * ld_out + r] = tile[tx][ty + 8 * i];
    }
}

int transpose_place(const void* in, void* out, int nb1, int nb2, int rows, int cols, int elem_bytes, long long stride_b1,
                    long long stride_b2, long long ld_out, int col_limit, cudaStream_t stream) {
    if (nb1 <= 0 || nb2 <= 0 || rows <= 0 || cols <= 0 || col_limit <= 0) return 0;
    if (elem_bytes != 4 && elem_bytes != 8) return fail("transpose: element size must be 4 or 8 bytes (got %d)", elem_bytes);
    const unsigned gy = (unsigned)((rows + 31) / 32);
    if ((long long)nb1 * nb2 > 65535 || gy > 65535) return fail("transpose: too many batches / rows per launch");
    if (ld_out < rows) return fail("transpose: output leading dimension %lld < rows %d", ld_out, rows);
    dim3 grid((cols + 31) / 32, gy, nb1 * nb2);
    if (elem_bytes == 4)
        transpose_place_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(in), static_cast<float*>(out), nb2,
                                                                rows, cols, stride_b1, stride_b2, ld_out, col_limit);
    else
        transpose_place_kernel<float2><<<grid, 256, 0, stream>>>(static_cast<const float2*>(in), static_cast<float2*>(out), nb2,
                                                                 rows, cols, stride_b1, stride_b2, ld_out, col_limit);
    SPYB_LAUNCH_CHECK("transpose_place_kernel");
    count_launch();
    return 0;
}

// ---------------------------------------------------------------------------------------
// row gather: dst[t][i][:] = src[t][idx[i]][:]   (time post-selection, compRoutines.py:593)
// ---------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx, int n_idx,
                                   long long row_elems, long long src_trial_stride, float* __restrict__ dst) {
    const int i = blockIdx.x, t = blockIdx.y;
    const float* __restrict__ s = src + (long long)t * src_trial_stride + (long long)idx[i] * row_elems;
    float* __restrict__ d = dst + ((long long)t * n_idx + i) * row_elems;
    for (long long e = threadIdx.x; e < row_elems; e += blockDim.x) d[e] = s[e];
}

int gather_rows(const float* src, int n_trials, long long src_trial_stride, const int* idx, int n_idx,
                long long row_elems, float* dst, cudaStream_t stream) {
    if (n_trials <= 0 || n_idx <= 0 || row_elems <= 0) return 0;
    if (n_trials > 65535) return fail("gather_rows: too many trials per launch");
    dim3 grid(n_idx, n_trials);
    gather_rows_kernel<<<grid, 256, 0, stream>>>(src, idx, n_idx, row_elems, src_trial_stride, dst);
    SPYB_LAUNCH_CHECK("gather_rows_kernel");
    count_launch();
    return 0;
}

}  // namespace spyb
