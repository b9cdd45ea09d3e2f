// Small element-wise kernels of the "next" rows of SURVEY 8(f): jackknife replicates / bias / variance
// (syncopy/statistics/jackknifing.py:14-184), pairwise phase consistency (ST_compRoutines.py:158-233 +
// connectivity_analysis.py:624-667) and the lag bookkeeping of the cross-covariance
// (ST_compRoutines.py:465-584).  All HBM-bound streaming kernels over [nFreq][C][C]-sized arrays.
#include "common.cuh"
#include "spyb_internal.h"

namespace spyb {
namespace {

inline unsigned grid_for(long long n, int per_block = 256) {
    long long b = (n + per_block - 1) / per_block;
    if (b > 148 * 32) b = 148 * 32;
    return (unsigned)(b < 1 ? 1 : b);
}

// out = a*x + b*y (float32 words; complex arrays count two words per element)
__global__ void axpby_kernel(const float* __restrict__ x, const float* __restrict__ y, float a, float b,
                             float* __restrict__ out, long long n) {
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
        out[i] = a * x[i] + (y ? b * y[i] : 0.f);
}

// var[e] += |avg[e] - x[e]|^2  (jackknifing.py:167-170); `words` = 2 for complex input (var is always real)
__global__ void sqdev_kernel(const float* __restrict__ avg, const float* __restrict__ x, float* __restrict__ var,
                             long long n_elem, int words) {
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_elem; i += step) {
        float d2;
        if (words == 2) {
            const float dr = avg[2 * i] - x[2 * i], di = avg[2 * i + 1] - x[2 * i + 1];
            const float m = hypotf(dr, di);             // np.abs of a complex64, then squared -- like the reference
            d2 = m * m;
        } else {
            const float d = fabsf(avg[i] - x[i]);
            d2 = d * d;
        }
        var[i] += d2;
    }
}

// acc[e] += z[e] / |z[e]|: unit phase vector of a single-trial cross spectrum (angle(0) = 0 like np.angle)
__global__ void unit_accumulate_kernel(const float2* __restrict__ z, float2* __restrict__ acc, long long n, int first) {
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const float2 v = z[i];
        const float m = hypotf(v.x, v.y);
        float2 u = m > 0.f ? make_float2(v.x / m, v.y / m) : make_float2(1.f, 0.f);
        if (!first) { const float2 a = acc[i]; u.x += a.x; u.y += a.y; }
        acc[i] = u;
    }
}

// ppc = (|sum_k u_k|^2 - T) / (T (T - 1)) = 2 / (T (T-1)) * sum_{j<k} cos(theta_j - theta_k)
__global__ void ppc_finish_kernel(const float2* __restrict__ acc, float* __restrict__ out, long long n, float T) {
    const long long step = (long long)gridDim.x * blockDim.x;
    const float inv = 1.f / (T * (T - 1.f));
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const float2 a = acc[i];
        out[i] = (a.x * a.x + a.y * a.y - T) * inv;
    }
}

// Cross-covariance, kernel spectra: channel j's time-reversed series h_j[t] = x_j[n-1-t] has the spectrum
// conj(X_j[k]) e^{-2 pi i k (n-1)/L}; an extra shift by `shift` samples makes the circular convolution come out as
// out[t] = full[t + shift].  T[j][k] = H_j[k] e^{+2 pi i k shift / L} / L for all L bins (conjugate symmetric).
__global__ void xcov_kern_kernel(const float2* __restrict__ X, int n_chan, int L, int n, int shift,
                                 float2* __restrict__ T) {
    const int j = blockIdx.y;
    const int half = L / 2;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < L; k += gridDim.x * blockDim.x) {
        const int kk = k <= half ? k : L - k;
        float2 x = X[(long long)j * (half + 1) + kk];          // [chan][bin]
        if (k > half) x.y = -x.y;                               // X[L-k] = conj X[k] (real input)
        // exponent (shift - (n-1)) * k mod L, reduced in integers before the sincos
        long long e = ((long long)(shift - (n - 1)) * k) % L;
        if (e < 0) e += L;
        float sn, cs;
        sincospif(2.0f * (float)e / (float)L, &sn, &cs);
        const float2 h = make_float2(x.x, -x.y);                // conj(X_j)
        const float inv = 1.0f / (float)L;
        T[(long long)j * L + k] = make_float2((h.x * cs - h.y * sn) * inv, (h.x * sn + h.y * cs) * inv);
    }
}

// corr[j][i][t] = full_ij[t + shift] (t < 2 nl + 1, shift = n - 1 - nl) -> CC[s][i][j] for the pair i >= j:
//   CC[s, i, j] = full[n-1+s] / (n - s),   CC[s, j, i] = full[n-1-s-e] / (n - s), e = 1 for even n
// (ST_compRoutines.py:555-566: 'same' slice of an even-length convolution is off-centre by one sample).
__global__ void xcov_finish_kernel(const float* __restrict__ corr, const float2* __restrict__ X, int n_chan, int n,
                                   int nl, int half1, int norm, float* __restrict__ out) {
    const int s = blockIdx.y;
    const int e = (n % 2 == 0) ? 1 : 0;
    const long long nt = 2LL * nl + 1;
    const long long pairs = (long long)n_chan * n_chan;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < pairs; q += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(q / n_chan), j = (int)(q % n_chan);
        float v;
        if (i >= j) v = corr[((long long)j * n_chan + i) * nt + (nl + s)];
        else        v = corr[((long long)i * n_chan + j) * nt + (nl - s - e)];     // pair (j, i) with j > i, mirrored lag
        v /= (float)(n - s);
        if (norm) {
            // np.std of the (detrended) series: sqrt(mean(x^2) - mean(x)^2) from the lag-0 autocorrelation and the DC bin
            const float vi = corr[((long long)i * n_chan + i) * nt + nl] / (float)n;
            const float vj = corr[((long long)j * n_chan + j) * nt + nl] / (float)n;
            const float mi = X[(long long)i * half1].x / (float)n, mj = X[(long long)j * half1].x / (float)n;
            v /= sqrtf(fmaxf(vi - mi * mi, 0.f)) * sqrtf(fmaxf(vj - mj * mj, 0.f));
        }
        out[(long long)s * pairs + q] = v;
    }
}

}  // namespace

int axpby(const float* x, const float* y, float a, float b, float* out, long long n, cudaStream_t st) {
    if (n <= 0) return 0;
    axpby_kernel<<<grid_for(n), 256, 0, st>>>(x, y, a, b, out, n);
    SPYB_LAUNCH_CHECK("axpby_kernel");
    count_launch();
    return 0;
}

int sqdev_accumulate(const float* avg, const float* x, float* var, long long n_elem, int is_complex, cudaStream_t st) {
    if (n_elem <= 0) return 0;
    sqdev_kernel<<<grid_for(n_elem), 256, 0, st>>>(avg, x, var, n_elem, is_complex ? 2 : 1);
    SPYB_LAUNCH_CHECK("sqdev_kernel");
    count_launch();
    return 0;
}

int unit_accumulate(const void* z, void* acc, long long n, int first, cudaStream_t st) {
    if (n <= 0) return 0;
    unit_accumulate_kernel<<<grid_for(n), 256, 0, st>>>(static_cast<const float2*>(z), static_cast<float2*>(acc), n, first);
    SPYB_LAUNCH_CHECK("unit_accumulate_kernel");
    count_launch();
    return 0;
}

int ppc_finish(const void* acc, float* out, long long n, int n_trials, cudaStream_t st) {
    if (n <= 0) return 0;
    if (n_trials < 2) return fail("ppc needs at least two trials (got %d)", n_trials);
    ppc_finish_kernel<<<grid_for(n), 256, 0, st>>>(static_cast<const float2*>(acc), out, n, (float)n_trials);
    SPYB_LAUNCH_CHECK("ppc_finish_kernel");
    count_launch();
    return 0;
}

int xcov_kernel_spectra(const void* xspec, int n_chan, int L, int n, int shift, void* kern, cudaStream_t st) {
    if (n_chan <= 0 || L <= 0) return 0;
    dim3 grid((L + 255) / 256, n_chan);
    xcov_kern_kernel<<<grid, 256, 0, st>>>(static_cast<const float2*>(xspec), n_chan, L, n, shift, static_cast<float2*>(kern));
    SPYB_LAUNCH_CHECK("xcov_kern_kernel");
    count_launch();
    return 0;
}

int xcov_finish(const float* corr, const void* xspec, int n_chan, int n, int n_lags, int L, int norm, float* out,
                cudaStream_t st) {
    if (n_chan <= 0 || n_lags <= 0) return 0;
    const long long pairs = (long long)n_chan * n_chan;
    dim3 grid(grid_for(pairs), n_lags);
    xcov_finish_kernel<<<grid, 256, 0, st>>>(corr, static_cast<const float2*>(xspec), n_chan, n, n_lags, L / 2 + 1, norm, out);
    SPYB_LAUNCH_CHECK("xcov_finish_kernel");
    count_launch();
    return 0;
}

}  // namespace spyb
