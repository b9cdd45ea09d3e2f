// Preprocessing compute functions: the sample-recursive and polyphase parts of
//   syncopy/preproc/compRoutines.py:175-276   but_filtering_cF   (scipy.signal.sosfilt / sosfiltfilt)
//   syncopy/preproc/compRoutines.py:541-616   resample_cF        (scipy.signal.resample_poly = upfirdn)
//   syncopy/preproc/compRoutines.py:765-832   standardize_cF,  :303-338 rectify_cF
// The FIR / Hilbert filters (sinc_filtering_cF, hilbert_cF) run as FFT convolutions on the wavelet kernel (cwt.cu).
// Arithmetic in float64 where the reference's is (SciPy promotes float32 data to the float64 of the coefficients).
#include "common.cuh"
#include "spyb_internal.h"

namespace spyb {
namespace {

constexpr int MAX_SECTIONS = 16;

struct SosArgs {
    const float* x;            // [trial][sample][channel]
    long long trial_stride;
    int n_trials, n_samples, n_chan;
    int n_sections;
    int edge;                  // odd extension on both sides (sosfiltfilt's padlen); 0 for a single forward pass
    int twopass;
    double sos[MAX_SECTIONS][6];
    double zi[MAX_SECTIONS][2];   // steady-state initial conditions (sosfilt_zi), scaled by the first sample; twopass only
    double* scratch;           // twopass: [trial][n_samples + 2 edge][channel] forward-filtered extension
    float* out;                // [trial][sample][channel]
};

// odd extension of scipy.signal._arraytools.odd_ext: 2 x[0] - x[edge - n] | x | 2 x[N-1] - x[N-2-k]
__device__ __forceinline__ double ext_value(const float* __restrict__ xc, int n_chan, int N, int edge, int n) {
    if (n < edge) return 2.0 * (double)xc[0] - (double)xc[(long long)(edge - n) * n_chan];
    if (n < edge + N) return (double)xc[(long long)(n - edge) * n_chan];
    const int k = n - edge - N;
    return 2.0 * (double)xc[(long long)(N - 1) * n_chan] - (double)xc[(long long)(N - 2 - k) * n_chan];
}

// One thread per (trial, channel): the recursion is sequential in time, parallel across channels and trials.
// Direct form II transposed, section after section per sample -- scipy.signal.sosfilt's loop.
__global__ void __launch_bounds__(128) sosfilt_kernel(const SosArgs a) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int trial = blockIdx.y;
    if (c >= a.n_chan) return;
    const float* __restrict__ xc = a.x + (long long)trial * a.trial_stride + c;
    const int N = a.n_samples, S = a.n_sections, edge = a.edge, next = N + 2 * edge;
    double z[MAX_SECTIONS][2];
    const double x0 = a.twopass ? ext_value(xc, a.n_chan, N, edge, 0) : 0.0;
#pragma unroll 1
    for (int s = 0; s < S; ++s) { z[s][0] = a.zi[s][0] * x0; z[s][1] = a.zi[s][1] * x0; }
    float* __restrict__ oc = a.out + (long long)trial * N * a.n_chan + c;
    double* __restrict__ sc = a.twopass ? a.scratch + (long long)trial * next * a.n_chan + c : nullptr;
    for (int n = 0; n < next; ++n) {
        double v = a.twopass ? ext_value(xc, a.n_chan, N, edge, n) : (double)xc[(long long)n * a.n_chan];
#pragma unroll 1
        for (int s = 0; s < S; ++s) {
            const double y = a.sos[s][0] * v + z[s][0];
            z[s][0] = a.sos[s][1] * v - a.sos[s][4] * y + z[s][1];
            z[s][1] = a.sos[s][2] * v - a.sos[s][5] * y;
            v = y;
        }
        if (a.twopass) sc[(long long)n * a.n_chan] = v;
        else oc[(long long)n * a.n_chan] = (float)v;
    }
    if (!a.twopass) return;
    // backward pass over the forward result, initial state scaled by its last sample (sosfiltfilt)
    const double y0 = sc[(long long)(next - 1) * a.n_chan];
#pragma unroll 1
    for (int s = 0; s < S; ++s) { z[s][0] = a.zi[s][0] * y0; z[s][1] = a.zi[s][1] * y0; }
    for (int n = next - 1; n >= 0; --n) {
        double v = sc[(long long)n * a.n_chan];
#pragma unroll 1
        for (int s = 0; s < S; ++s) {
            const double y = a.sos[s][0] * v + z[s][0];
            z[s][0] = a.sos[s][1] * v - a.sos[s][4] * y + z[s][1];
            z[s][1] = a.sos[s][2] * v - a.sos[s][5] * y;
            v = y;
        }
        if (n >= edge && n < edge + N) oc[(long long)(n - edge) * a.n_chan] = (float)v;
    }
}

// y[m][c] = sum_i x[i][c] h[(m + m0) down - i up]: scipy.signal.upfirdn restricted to the rows resample_poly keeps
__global__ void __launch_bounds__(256) upfirdn_kernel(const float* __restrict__ x, long long trial_stride, int n_in,
                                                      int n_chan, const double* __restrict__ h, int len_h, int up,
                                                      int down, int m0, int n_out, float* __restrict__ out) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int m = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int trial = blockIdx.z;
    if (c >= n_chan || m >= n_out) return;
    const float* __restrict__ xt = x + (long long)trial * trial_stride + c;
    const long long t = (long long)(m + m0) * down;               // position in the upsampled series
    long long i_hi = t / up;                                      // h index >= 0
    if (i_hi > n_in - 1) i_hi = n_in - 1;
    long long i_lo = (t - len_h + 1 + up - 1) / up;               // h index <= len_h - 1
    if (t - len_h + 1 < 0) i_lo = 0;
    double acc = 0.0;
    for (long long i = i_lo; i <= i_hi; ++i) acc += (double)xt[i * n_chan] * h[t - i * up];
    out[((long long)trial * n_out + m) * n_chan + c] = (float)acc;
}

// (x - mean) / std per channel in float32 (np.mean / np.std over time, population std); block = 32 channels x 8 lanes
__global__ void __launch_bounds__(256) standardize_kernel(const float* __restrict__ x, long long trial_stride,
                                                          int n_samples, int n_chan, float* __restrict__ out) {
    __shared__ float red[8][33];
    const int cl = threadIdx.x & 31, tl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const float* __restrict__ xt = x + (long long)blockIdx.y * trial_stride;
    float* __restrict__ ot = out + (long long)blockIdx.y * n_samples * n_chan;
    float s = 0.f;
    if (c < n_chan) for (int n = tl; n < n_samples; n += 8) s += xt[(long long)n * n_chan + c];
    red[tl][cl] = s;
    __syncthreads();
    float mean = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) mean += red[t][cl];
    mean /= (float)n_samples;
    __syncthreads();
    float q = 0.f;
    if (c < n_chan) for (int n = tl; n < n_samples; n += 8) { const float d = xt[(long long)n * n_chan + c] - mean; q += d * d; }
    red[tl][cl] = q;
    __syncthreads();
    float var = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) var += red[t][cl];
    const float inv = 1.f / sqrtf(var / (float)n_samples);
    if (c < n_chan) for (int n = tl; n < n_samples; n += 8) {
        const long long o = (long long)n * n_chan + c;
        ot[o] = (xt[o] - mean) * inv;
    }
}

__global__ void abs_kernel(const float* __restrict__ x, float* __restrict__ out, long long n) {
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) out[i] = fabsf(x[i]);
}

}  // namespace

int sosfilt(const float* x, int n_trials, long long trial_stride, int n_samples, int n_chan, const double* sos_host,
            int n_sections, const double* zi_host, int edge, int twopass, double* scratch, float* out, cudaStream_t st) {
    if (n_trials <= 0 || n_samples <= 0 || n_chan <= 0) return 0;
    if (n_sections < 1 || n_sections > MAX_SECTIONS)
        return fail("sosfilt: 1..%d second-order sections supported (got %d)", MAX_SECTIONS, n_sections);
    if (n_trials > 65535) return fail("sosfilt: at most 65535 trials per call");
    if (twopass && (edge < 0 || edge >= n_samples))
        return fail("sosfiltfilt: the length of the input (%d) must be greater than padlen (%d)", n_samples, edge);
    if (twopass && (!scratch || !zi_host)) return fail("sosfiltfilt needs scratch memory and initial conditions");
    SosArgs a;
    a.x = x; a.trial_stride = trial_stride; a.n_trials = n_trials; a.n_samples = n_samples; a.n_chan = n_chan;
    a.n_sections = n_sections; a.edge = twopass ? edge : 0; a.twopass = twopass ? 1 : 0;
    for (int s = 0; s < n_sections; ++s) {
        for (int k = 0; k < 6; ++k) a.sos[s][k] = sos_host[s * 6 + k];
        a.zi[s][0] = zi_host ? zi_host[s * 2] : 0.0;
        a.zi[s][1] = zi_host ? zi_host[s * 2 + 1] : 0.0;
    }
    a.scratch = scratch; a.out = out;
    dim3 grid((n_chan + 127) / 128, n_trials);
    sosfilt_kernel<<<grid, 128, 0, st>>>(a);
    SPYB_LAUNCH_CHECK("sosfilt_kernel");
    count_launch();
    return 0;
}

int upfirdn(const float* x, int n_trials, long long trial_stride, int n_in, int n_chan, const double* h, int len_h,
            int up, int down, int m0, int n_out, float* out, cudaStream_t st) {
    if (n_trials <= 0 || n_out <= 0 || n_chan <= 0) return 0;
    if (up < 1 || down < 1 || len_h < 1) return fail("upfirdn: up, down and the filter length must be positive");
    if (n_trials > 65535 || (n_out + 7) / 8 > 65535) return fail("upfirdn: too many trials / output samples for one launch");
    dim3 grid((n_chan + 31) / 32, (n_out + 7) / 8, n_trials);
    upfirdn_kernel<<<grid, 256, 0, st>>>(x, trial_stride, n_in, n_chan, h, len_h, up, down, m0, n_out, out);
    SPYB_LAUNCH_CHECK("upfirdn_kernel");
    count_launch();
    return 0;
}

int standardize(const float* x, int n_trials, long long trial_stride, int n_samples, int n_chan, float* out, cudaStream_t st) {
    if (n_trials <= 0 || n_samples <= 0 || n_chan <= 0) return 0;
    if (n_trials > 65535) return fail("standardize: at most 65535 trials per call");
    dim3 grid((n_chan + 31) / 32, n_trials);
    standardize_kernel<<<grid, 256, 0, st>>>(x, trial_stride, n_samples, n_chan, out);
    SPYB_LAUNCH_CHECK("standardize_kernel");
    count_launch();
    return 0;
}

int rectify(const float* x, float* out, long long n, cudaStream_t st) {
    if (n <= 0) return 0;
    long long blocks = (n + 1023) / 1024;
    if (blocks > 148 * 16) blocks = 148 * 16;
    abs_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, out, n);
    SPYB_LAUNCH_CHECK("abs_kernel");
    count_launch();
    return 0;
}

}  // namespace spyb
