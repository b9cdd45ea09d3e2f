// Long transforms: FFT lengths beyond the shared-memory engines (power-of-two lengths > 16384, other lengths > 8192).
//
// The reference takes any trial length (scipy.fft.rfft in syncopy/specest/mtmfft.py:117-127, fftconvolve in
// syncopy/specest/wavelets/transform.py:88-108); the shared-memory kernels (mtm*.cu, cwt.cu) hold a whole transform
// per block and stop at 2^14 points.  Beyond that the data lives in global memory as [len][E] complex64 -- E =
// independent columns (channel pairs x tapers, or scales x channels) -- and the FFT runs along axis 0 as Stockham
// auto-sort passes with one thread per (butterfly, column): every load and store of a pass is contiguous across
// the threads of a warp, a pass moves the array once through HBM (radix 16 where the length allows: 4 passes for
// 65536 points).  Lengths with prime factors up to 61 run as mixed-radix passes, anything else through Bluestein's
// chirp-z on a power-of-two length.  The tapered-FFT front end (detrend, taper, pair packing) and back end (pair
// split, scale, frequency gather, output conversion, taper mean) are the same operations as in mtm.cu, as
// separate HBM-bound kernels.
#include "common.cuh"
#include "fft_core.cuh"
#include "spyb_internal.h"

#include <vector>

namespace spyb {
namespace {

constexpr int LT = 128;                      // threads per block
constexpr int MAX_GENERIC_RADIX = 61;        // larger prime factors -> Bluestein
constexpr long long MAX_LONG_LEN = 1LL << 24;
constexpr size_t WORK_BUDGET = (size_t)1 << 30;   // bytes per ping-pong buffer the drivers aim for

inline unsigned blocks_for(long long n) { return (unsigned)((n + LT - 1) / LT); }

// ---------------------------------------------------------------------------------------------------------
// tables
// ---------------------------------------------------------------------------------------------------------
__global__ void ltwiddle_kernel(float2* tw, int len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    double s, c;
    sincospi(-2.0 * (double)i / (double)len, &s, &c);
    tw[i] = make_float2((float)c, (float)s);
}

// b[i] = e^{+i pi i^2 / n} (phase reduced exactly: i^2 mod 2n in 64-bit integers); bw = b wrapped onto length M
__global__ void lchirp_kernel(float2* chirp, float2* bw, int n, int M) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    float2 wrapped = make_float2(0.f, 0.f);
    const int src = i < n ? i : (M - i < n ? M - i : -1);
    if (src >= 0) {
        const long long q = ((long long)src * src) % (2LL * n);
        double s, c;
        sincospi((double)q / (double)n, &s, &c);
        wrapped = make_float2((float)c, (float)s);
        if (i < n) chirp[i] = wrapped;
    }
    bw[i] = wrapped;
}

// ---------------------------------------------------------------------------------------------------------
// Stockham passes over [len][E]
// ---------------------------------------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void ldft(float2 (&x)[R], const float2* __restrict__ tw, int len) {
    if constexpr (R == 2 || R == 4 || R == 8 || R == 16) {
        Radix<R>::run(x);
        if constexpr (R >= 8) {
            float2 y[R];
#pragma unroll
            for (int k = 0; k < R; ++k) y[k] = x[Radix<R>::reg_of(k)];
#pragma unroll
            for (int k = 0; k < R; ++k) x[k] = y[k];
        }
    } else {
        // small odd DFT straight from the twiddle table: W_R^m = tw[m * len / R]
        float2 y[R];
        const int step = len / R;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            float2 s = x[0];
#pragma unroll
            for (int t = 1; t < R; ++t) {
                const float2 w = __ldg(tw + ((k * t) % R) * step);
                s.x += x[t].x * w.x - x[t].y * w.y;
                s.y += x[t].x * w.y + x[t].y * w.x;
            }
            y[k] = s;
        }
#pragma unroll
        for (int k = 0; k < R; ++k) x[k] = y[k];
    }
}

// flat thread index g = j * E + e (butterfly j, column e): in[(j + t * len/R) * E + e] = in[g + t * (len/R) * E]
template <int R>
__global__ void __launch_bounds__(LT) lfft_pass_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                       const float2* __restrict__ tw, int len, int ns, int E) {
    const int stride = len / R;
    const long long total = (long long)stride * E;
    const long long g = (long long)blockIdx.x * LT + threadIdx.x;
    if (g >= total) return;
    const int j = (int)(g / E);
    const int e = (int)(g - (long long)j * E);
    const int k = j % ns;
    const int tstep = len / (ns * R);
    float2 x[R];
#pragma unroll
    for (int t = 0; t < R; ++t) x[t] = in[g + (long long)t * total];
    if (ns > 1) {
#pragma unroll
        for (int t = 1; t < R; ++t) x[t] = cmul(x[t], __ldg(tw + (long long)t * k * tstep));
    }
    ldft<R>(x, tw, len);
    const long long j0 = (long long)(j - k) * R + k;
#pragma unroll
    for (int q = 0; q < R; ++q) out[(j0 + (long long)q * ns) * E + e] = x[q];
}

// any prime radix: one output per thread, O(R) inputs each
__global__ void __launch_bounds__(LT) lfft_pass_generic_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                               const float2* __restrict__ tw, int len, int ns, int R,
                                                               int E) {
    const long long total = (long long)len * E;
    const long long g = (long long)blockIdx.x * LT + threadIdx.x;
    if (g >= total) return;
    const int jq = (int)(g / E);
    const int e = (int)(g - (long long)jq * E);
    const int j = jq / R, q = jq - j * R;
    const int k = j % ns;
    const int stride = len / R;
    const int tstep = len / (ns * R);
    float2 s = make_float2(0.f, 0.f);
    for (int t = 0; t < R; ++t) {
        const float2 v = in[((long long)j + (long long)t * stride) * E + e];
        const float2 w1 = __ldg(tw + (long long)t * k * tstep);
        const float2 w2 = __ldg(tw + (long long)((q * t) % R) * stride);
        const float2 vw = cmul(v, w1);
        s.x += vw.x * w2.x - vw.y * w2.y;
        s.y += vw.x * w2.y + vw.y * w2.x;
    }
    const long long j0 = (long long)(j - k) * R + k;
    out[(j0 + (long long)q * ns) * E + e] = s;
}

struct LongPlan {
    int len = 0;
    std::vector<int> radices;
    int max_prime = 1;
};

LongPlan factorize(int len) {
    LongPlan p;
    p.len = len;
    int r = len;
    while (r % 16 == 0) { p.radices.push_back(16); r /= 16; }
    if (r % 8 == 0) { p.radices.push_back(8); r /= 8; }
    if (r % 4 == 0) { p.radices.push_back(4); r /= 4; }
    if (r % 2 == 0) { p.radices.push_back(2); r /= 2; }
    for (int f = 3; r > 1; f += 2) {
        while (r % f == 0) { p.radices.push_back(f); r /= f; if (f > p.max_prime) p.max_prime = f; }
        if ((long long)f * f > r && r > 1) { p.radices.push_back(r); if (r > p.max_prime) p.max_prime = r; r = 1; }
    }
    return p;
}

template <int R>
int launch_lpass(const float2* in, float2* out, const float2* tw, int len, int ns, int E, cudaStream_t st) {
    const long long total = (long long)(len / R) * E;
    lfft_pass_kernel<R><<<blocks_for(total), LT, 0, st>>>(in, out, tw, len, ns, E);
    SPYB_LAUNCH_CHECK("lfft_pass_kernel");
    count_launch();
    return 0;
}

// forward FFT along axis 0 of [len][E]; data starts in `a`, the result pointer comes back through *res (a or b)
int lfft_axis0(float2* a, float2* b, const float2* tw, const LongPlan& plan, int E, float2** res, cudaStream_t st) {
    float2* src = a;
    float2* dst = b;
    const int len = plan.len;
    int ns = 1;
    if ((long long)len * E / 2 > (long long)LT * 0x7fffffffLL) return fail("long FFT: %d x %d does not fit one launch", len, E);
    for (int R : plan.radices) {
        int rc;
        switch (R) {
            case 16: rc = launch_lpass<16>(src, dst, tw, len, ns, E, st); break;
            case 8:  rc = launch_lpass<8>(src, dst, tw, len, ns, E, st); break;
            case 4:  rc = launch_lpass<4>(src, dst, tw, len, ns, E, st); break;
            case 2:  rc = launch_lpass<2>(src, dst, tw, len, ns, E, st); break;
            case 3:  rc = launch_lpass<3>(src, dst, tw, len, ns, E, st); break;
            case 5:  rc = launch_lpass<5>(src, dst, tw, len, ns, E, st); break;
            case 7:  rc = launch_lpass<7>(src, dst, tw, len, ns, E, st); break;
            default: {
                lfft_pass_generic_kernel<<<blocks_for((long long)len * E), LT, 0, st>>>(src, dst, tw, len, ns, R, E);
                SPYB_LAUNCH_CHECK("lfft_pass_generic_kernel");
                count_launch();
                rc = 0;
            }
        }
        if (rc) return rc;
        ns *= R;
        float2* t = src; src = dst; dst = t;
    }
    *res = src;
    return 0;
}

// stream-ordered scratch memory of one call
struct Scratch {
    cudaStream_t st;
    std::vector<void*> ptrs;
    explicit Scratch(cudaStream_t s) : st(s) {}
    ~Scratch() { for (void* p : ptrs) cudaFreeAsync(p, st); }
    template <typename T> T* take(size_t count) {
        void* p = nullptr;
        if (cudaMallocAsync(&p, count * sizeof(T) + 16, st) != cudaSuccess) {
            fail("long transform: cannot allocate %zu bytes of scratch memory", count * sizeof(T));
            cudaGetLastError();
            return nullptr;
        }
        ptrs.push_back(p);
        return static_cast<T*>(p);
    }
};

// ---------------------------------------------------------------------------------------------------------
// tapered-FFT front end
// ---------------------------------------------------------------------------------------------------------
constexpr int NSEG = 64;       // time segments of the detrend / taper-mean reductions (fixed order: deterministic)

// MODE 0: partial[seg][0][c] = sum x, partial[seg][1][c] = sum (n - tmid) x over the segment's samples of the window
// MODE 1: partial[(k * NSEG + seg)][0][c] = sum (x - trend) w_k[n]  (blockIdx.z = taper k)
template <int MODE>
__global__ void __launch_bounds__(256) lstats_kernel(const float* __restrict__ xt, long long start, int n_samples,
                                                     int n_chan, int c_begin, int c_count, int n_win,
                                                     const float* __restrict__ tapers, const double* __restrict__ trend,
                                                     double* __restrict__ partial) {
    __shared__ double red[2][8][33];
    const int cl = threadIdx.x & 31, tl = threadIdx.x >> 5;
    const int cc = blockIdx.x * 32 + cl;                      // channel within the chunk
    const int seg = blockIdx.y, k = blockIdx.z;
    const int per = (n_win + NSEG - 1) / NSEG;
    const int n0 = seg * per, n1 = min(n_win, n0 + per);
    const double tmid = 0.5 * (double)(n_win - 1);
    double s0 = 0.0, s1 = 0.0;
    if (cc < c_count) {
        const int c = c_begin + cc;
        double mean = 0.0, slope = 0.0;
        const float* __restrict__ win = nullptr;
        if (MODE == 1) { mean = trend[cc]; slope = trend[c_count + cc]; win = tapers + (long long)k * n_win; }
        for (int n = n0 + tl; n < n1; n += 8) {
            const long long m = start + n;
            const float v = (m >= 0 && m < n_samples) ? __ldg(xt + m * n_chan + c) : 0.f;
            if (MODE == 0) {
                s0 += (double)v;
                s1 += ((double)n - tmid) * (double)v;
            } else {
                const float d = v - (float)(mean + slope * ((double)n - tmid));
                s0 += (double)(d * __ldg(win + n));
            }
        }
    }
    red[0][tl][cl] = s0; red[1][tl][cl] = s1;
    __syncthreads();
    if (tl == 0 && cc < c_count) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int t = 0; t < 8; ++t) { a0 += red[0][t][cl]; a1 += red[1][t][cl]; }
        double* dst = partial + ((long long)(k * NSEG + seg) * 2) * c_count;
        dst[cc] = a0;
        dst[c_count + cc] = a1;
    }
}

// trend[0][c] = mean, trend[1][c] = slope (MODE 0); tmean[k][c] = sum / n_win (MODE 1)
__global__ void lstats_finish_kernel(const double* __restrict__ partial, int c_count, int n_win, int polyremoval,
                                     int mode, double* __restrict__ outv) {
    const int cc = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (cc >= c_count) return;
    double a0 = 0.0, a1 = 0.0;
    for (int seg = 0; seg < NSEG; ++seg) {
        const double* src = partial + ((long long)(k * NSEG + seg) * 2) * c_count;
        a0 += src[cc]; a1 += src[c_count + cc];
    }
    if (mode == 0) {
        double mean = a0 / (double)n_win, slope = 0.0;
        if (polyremoval == 1 && n_win > 1) slope = a1 / ((double)n_win * ((double)n_win * n_win - 1.0) / 12.0);
        if (polyremoval < 0) mean = 0.0;
        outv[cc] = mean;
        outv[c_count + cc] = slope;
    } else {
        outv[(long long)k * c_count + cc] = a0 / (double)n_win;
    }
}

// A[n][e] (e = k * n_pairs + p) = tapered, detrended samples of channel pair p packed as one complex series; rows
// n >= n_win are the zero padding up to the transform length; Bluestein: times conj(b[n])
__global__ void __launch_bounds__(LT) lpack_kernel(const float* __restrict__ xt, long long start, int n_samples,
                                                   int n_chan, int c_begin, int c_count, int n_pairs, int n_win,
                                                   int n_fft, const float* __restrict__ tapers, int n_tapers,
                                                   const double* __restrict__ trend, const double* __restrict__ tmean,
                                                   const float2* __restrict__ chirp, float2* __restrict__ A) {
    const int E = n_tapers * n_pairs;
    const long long total = (long long)n_fft * E;
    const long long g = (long long)blockIdx.x * LT + threadIdx.x;
    if (g >= total) return;
    const int n = (int)(g / E);
    const int e = (int)(g - (long long)n * E);
    float2 v = make_float2(0.f, 0.f);
    if (n < n_win) {
        const int k = e / n_pairs, p = e - k * n_pairs;
        const int ca = 2 * p, cb = 2 * p + 1;                 // within the chunk
        const long long m = start + n;
        const float w = __ldg(tapers + (long long)k * n_win + n);
        const double t = (double)n - 0.5 * (double)(n_win - 1);
        const bool in = m >= 0 && m < n_samples;
        {
            const float x = in ? __ldg(xt + m * n_chan + c_begin + ca) : 0.f;
            v.x = (x - (float)(trend[ca] + trend[c_count + ca] * t)) * w;
            if (tmean) v.x -= (float)tmean[(long long)k * c_count + ca];
        }
        if (cb < c_count) {
            const float x = in ? __ldg(xt + m * n_chan + c_begin + cb) : 0.f;
            v.y = (x - (float)(trend[cb] + trend[c_count + cb] * t)) * w;
            if (tmean) v.y -= (float)tmean[(long long)k * c_count + cb];
        }
        if (chirp) v = cmulc(v, __ldg(chirp + n));
    }
    A[g] = v;
}

// Bluestein middle step: V = conj(FFT(a) * bhat) / M, in place (the second forward FFT then yields conj of the
// inverse transform)
__global__ void __launch_bounds__(LT) lblue_mul_kernel(float2* __restrict__ V, const float2* __restrict__ bhat, int M,
                                                       int E, float inv_m) {
    const long long total = (long long)M * E;
    const long long g = (long long)blockIdx.x * LT + threadIdx.x;
    if (g >= total) return;
    const int k = (int)(g / E);
    const float2 y = cmul(V[g], __ldg(bhat + k));
    V[g] = make_float2(y.x * inv_m, -y.y * inv_m);
}

struct FinishArgs {
    const float2* F;           // [n_fft][E]
    const float2* chirp;       // Bluestein: Z[k] = conj(b[k] * F[k]); else null
    int n_fft, n_dft, E, n_pairs, n_tapers;
    int c_begin, c_count;
    const int* freq_idx;
    int n_freq_out;
    int out_kind, keeptapers;
    float half_scale;
    void* out;
    long long off0, so_taper, so_freq;   // off0 = trial * so_trial + frame * so_frame
    int n_chan;
    float* chan_amax;
};

__global__ void __launch_bounds__(LT) lfinish_kernel(const FinishArgs a) {
    const long long total = (long long)a.n_freq_out * a.n_pairs;
    const long long g = (long long)blockIdx.x * LT + threadIdx.x;
    if (g >= total) return;
    const int fi = (int)(g / a.n_pairs);
    const int p = (int)(g - (long long)fi * a.n_pairs);
    const int kf = a.freq_idx ? __ldg(a.freq_idx + fi) : fi;
    const int kn = kf == 0 ? 0 : a.n_dft - kf;
    const int ca = 2 * p;
    const bool cb_ok = ca + 1 < a.c_count;
    const int c = a.c_begin + ca;
    float2 wf = make_float2(1.f, 0.f), wn = wf;
    if (a.chirp) { wf = __ldg(a.chirp + kf); wn = __ldg(a.chirp + kn); }
    const float inv_ntap = 1.f / (float)a.n_tapers;
    float2 acc_a = make_float2(0.f, 0.f), acc_b = acc_a;
    float amax_a = 0.f, amax_b = 0.f;
    for (int k = 0; k < a.n_tapers; ++k) {
        const int e = k * a.n_pairs + p;
        float2 z1 = a.F[(long long)kf * a.E + e];
        float2 z2 = a.F[(long long)kn * a.E + e];
        if (a.chirp) { z1 = cconj(cmul(z1, wf)); z2 = cconj(cmul(z2, wn)); }
        const float2 xa = make_float2((z1.x + z2.x) * a.half_scale, (z1.y - z2.y) * a.half_scale);
        const float2 xb = make_float2((z1.y + z2.y) * a.half_scale, (z2.x - z1.x) * a.half_scale);
        amax_a = fmaxf(amax_a, fmaxf(fabsf(xa.x), fabsf(xa.y)));
        amax_b = fmaxf(amax_b, fmaxf(fabsf(xb.x), fabsf(xb.y)));
        if (a.keeptapers) {
            const long long off = a.off0 + (long long)k * a.so_taper + (long long)fi * a.so_freq + c;
            if (a.out_kind == OUT_FOURIER_PLANAR) {
                float* o = reinterpret_cast<float*>(a.out) + off;
                o[0] = xa.x; o[a.n_chan] = xa.y;
                if (cb_ok) { o[1] = xb.x; o[a.n_chan + 1] = xb.y; }
            } else if (a.out_kind == OUT_FOURIER) {
                float2* o = reinterpret_cast<float2*>(a.out) + off;
                o[0] = xa;
                if (cb_ok) o[1] = xb;
            } else {
                float* o = reinterpret_cast<float*>(a.out) + off;
                o[0] = convert_real(xa, a.out_kind);
                if (cb_ok) o[1] = convert_real(xb, a.out_kind);
            }
        } else if (a.out_kind == OUT_FOURIER) {
            acc_a.x += xa.x; acc_a.y += xa.y; acc_b.x += xb.x; acc_b.y += xb.y;
        } else {
            acc_a.x += convert_real(xa, a.out_kind);
            acc_b.x += convert_real(xb, a.out_kind);
        }
    }
    if (!a.keeptapers) {
        const long long off = a.off0 + (long long)fi * a.so_freq + c;
        const float s = a.n_tapers > 1 ? inv_ntap : 1.f;
        if (a.out_kind == OUT_FOURIER) {
            float2* o = reinterpret_cast<float2*>(a.out) + off;
            o[0] = make_float2(acc_a.x * s, acc_a.y * s);
            if (cb_ok) o[1] = make_float2(acc_b.x * s, acc_b.y * s);
        } else {
            float* o = reinterpret_cast<float*>(a.out) + off;
            o[0] = acc_a.x * s;
            if (cb_ok) o[1] = acc_b.x * s;
        }
    }
    if (a.chan_amax) {
        atomicMax(reinterpret_cast<int*>(a.chan_amax) + c, __float_as_int(amax_a));
        if (cb_ok) atomicMax(reinterpret_cast<int*>(a.chan_amax) + c + 1, __float_as_int(amax_b));
    }
}

struct BlueTables {
    const float2* chirp = nullptr;   // [n]
    const float2* bhat = nullptr;    // [M] = FFT_M(b wrapped), not yet divided by M
};

// chirp tables of a Bluestein transform of length n on the power-of-two length M, built on the device
int make_blue_tables(int n, int M, const LongPlan& plan_m, const float2* tw_m, Scratch& sc, BlueTables* out,
                     cudaStream_t st) {
    float2* chirp = sc.take<float2>(n);
    float2* bw = sc.take<float2>(M);
    float2* bw2 = sc.take<float2>(M);
    if (!chirp || !bw || !bw2) return 1;
    lchirp_kernel<<<blocks_for(M), LT, 0, st>>>(chirp, bw, n, M);
    SPYB_LAUNCH_CHECK("lchirp_kernel");
    count_launch();
    float2* res = nullptr;
    if (lfft_axis0(bw, bw2, tw_m, plan_m, 1, &res, st)) return 1;
    out->chirp = chirp;
    out->bhat = res;
    return 0;
}

}  // namespace

bool mtm_needs_long(int n_dft) {
    const bool pow2 = (n_dft & (n_dft - 1)) == 0;
    return pow2 ? n_dft > 16384 : n_dft > 8192;
}

// Tapered real FFT of frames for transform lengths beyond the shared-memory kernels; same contract as mtm_frames.
int mtm_frames_long(const MtmFramesDesc& d, cudaStream_t st) {
    if (d.n_dft > MAX_LONG_LEN) return fail("FFT length %d exceeds the supported maximum of %lld", d.n_dft, MAX_LONG_LEN);
    if (d.out_kind == OUT_FOURIER_PLANAR && !d.keeptapers) return fail("planar complex output needs keeptapers = 1");
    const int L = d.n_dft;
    LongPlan plan = factorize(L);
    const bool blue = plan.max_prime > MAX_GENERIC_RADIX;
    int n_fft = L;
    if (blue) {
        long long m = 16;
        while (m < 2LL * L - 1) m <<= 1;
        if (m > 2 * MAX_LONG_LEN) return fail("FFT length %d is not supported", L);
        n_fft = (int)m;
        plan = factorize(n_fft);
    }
    Scratch sc(st);
    float2* tw = sc.take<float2>(n_fft);
    if (!tw) return 1;
    ltwiddle_kernel<<<blocks_for(n_fft), LT, 0, st>>>(tw, n_fft);
    SPYB_LAUNCH_CHECK("ltwiddle_kernel");
    count_launch();
    BlueTables bt;
    if (blue && make_blue_tables(L, n_fft, plan, tw, sc, &bt, st)) return 1;

    const int K = d.n_tapers;
    const int pairs_all = (d.n_chan + 1) / 2;
    long long pc = (long long)(WORK_BUDGET / ((size_t)n_fft * K * sizeof(float2)));
    if (pc < 1) pc = 1;
    if (pc > pairs_all) pc = pairs_all;
    const int max_cc = (int)(2 * pc);
    float2* bufA = sc.take<float2>((size_t)n_fft * K * pc);
    float2* bufB = sc.take<float2>((size_t)n_fft * K * pc);
    double* partial = sc.take<double>((size_t)K * NSEG * 2 * max_cc);
    double* trend = sc.take<double>((size_t)2 * max_cc);
    double* tmean = d.demean_taper ? sc.take<double>((size_t)K * max_cc) : nullptr;
    if (!bufA || !bufB || !partial || !trend || (d.demean_taper && !tmean)) return 1;

    for (int trial = 0; trial < d.n_trials; ++trial) {
        const float* xt = d.x + (long long)trial * d.trial_stride;
        for (int frame = 0; frame < d.n_frames; ++frame) {
            const long long start = (long long)d.frame_start0 + (long long)frame * d.hop;
            for (int p0 = 0; p0 < pairs_all; p0 += (int)pc) {
                const int np = (int)(p0 + pc <= pairs_all ? pc : pairs_all - p0);
                const int c_begin = 2 * p0;
                const int c_count = (c_begin + 2 * np <= d.n_chan) ? 2 * np : d.n_chan - c_begin;
                const int E = K * np;
                const dim3 sgrid((c_count + 31) / 32, NSEG, 1);
                if (d.polyremoval >= 0) {
                    lstats_kernel<0><<<sgrid, 256, 0, st>>>(xt, start, d.n_samples, d.n_chan, c_begin, c_count, d.n_win,
                                                           nullptr, nullptr, partial);
                    SPYB_LAUNCH_CHECK("lstats_kernel");
                    count_launch();
                    lstats_finish_kernel<<<dim3((c_count + 127) / 128, 1), 128, 0, st>>>(partial, c_count, d.n_win,
                                                                                         d.polyremoval, 0, trend);
                    SPYB_LAUNCH_CHECK("lstats_finish_kernel");
                    count_launch();
                } else {
                    SPYB_CUDA(cudaMemsetAsync(trend, 0, sizeof(double) * 2 * c_count, st));
                }
                if (d.demean_taper) {
                    lstats_kernel<1><<<dim3(sgrid.x, NSEG, K), 256, 0, st>>>(xt, start, d.n_samples, d.n_chan, c_begin,
                                                                             c_count, d.n_win, d.tapers, trend, partial);
                    SPYB_LAUNCH_CHECK("lstats_kernel");
                    count_launch();
                    lstats_finish_kernel<<<dim3((c_count + 127) / 128, K), 128, 0, st>>>(partial, c_count, d.n_win,
                                                                                         d.polyremoval, 1, tmean);
                    SPYB_LAUNCH_CHECK("lstats_finish_kernel");
                    count_launch();
                }
                lpack_kernel<<<blocks_for((long long)n_fft * E), LT, 0, st>>>(
                    xt, start, d.n_samples, d.n_chan, c_begin, c_count, np, d.n_win, n_fft, d.tapers, K, trend, tmean,
                    blue ? bt.chirp : nullptr, bufA);
                SPYB_LAUNCH_CHECK("lpack_kernel");
                count_launch();
                float2* res = nullptr;
                if (lfft_axis0(bufA, bufB, tw, plan, E, &res, st)) return 1;
                if (blue) {
                    lblue_mul_kernel<<<blocks_for((long long)n_fft * E), LT, 0, st>>>(res, bt.bhat, n_fft, E,
                                                                                      1.f / (float)n_fft);
                    SPYB_LAUNCH_CHECK("lblue_mul_kernel");
                    count_launch();
                    float2* other = res == bufA ? bufB : bufA;
                    float2* res2 = nullptr;
                    if (lfft_axis0(res, other, tw, plan, E, &res2, st)) return 1;
                    res = res2;
                }
                FinishArgs fa;
                fa.F = res; fa.chirp = blue ? bt.chirp : nullptr;
                fa.n_fft = n_fft; fa.n_dft = L; fa.E = E; fa.n_pairs = np; fa.n_tapers = K;
                fa.c_begin = c_begin; fa.c_count = c_count;
                fa.freq_idx = d.freq_idx; fa.n_freq_out = d.n_freq_out;
                fa.out_kind = d.out_kind; fa.keeptapers = d.keeptapers;
                fa.half_scale = 0.5f * d.scale;
                fa.out = d.out;
                fa.off0 = (long long)trial * d.so_trial + (long long)frame * d.so_frame;
                fa.so_taper = d.so_taper; fa.so_freq = d.so_freq;
                fa.n_chan = d.n_chan;
                fa.chan_amax = d.chan_amax;
                lfinish_kernel<<<blocks_for((long long)d.n_freq_out * np), LT, 0, st>>>(fa);
                SPYB_LAUNCH_CHECK("lfinish_kernel");
                count_launch();
            }
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// wavelet / superlet transforms with a circular length beyond the shared-memory kernel (cwt.cu)
// ---------------------------------------------------------------------------------------------------------
namespace {

struct CwtLongArgs {
    const float2* X;           // spectra of one trial: [L/2+1][chan] (or [chan][L/2+1] when transposed)
    int transposed;
    const float2* kern;        // [scale][max_fac][L]
    const float* expo;
    const int* n_fac;
    int n_chan, L, max_fac, n_scales;
    int s_begin, s_count;      // scales of this chunk; column e = s_local * n_chan + c
    int fj;
    int n_time, out_kind;
    float2* zacc;              // [n_time][E] running product (or log-magnitude in .x)
    void* out;                 // trial base; [n_time][scale][chan] (or [scale][chan][n_time] when transposed)
};

// V[k][e] = conj(X_c[k] T_{s,fj}[k]) over the full circle (X extended by conjugate symmetry)
__global__ void __launch_bounds__(LT) lcwt_mul_kernel(const CwtLongArgs a, float2* __restrict__ V) {
    const int E = a.s_count * a.n_chan;
    const long long total = (long long)a.L * E;
    const long long g = (long long)blockIdx.x * LT + threadIdx.x;
    if (g >= total) return;
    const int k = (int)(g / E);
    const int e = (int)(g - (long long)k * E);
    const int sl = e / a.n_chan, c = e - sl * a.n_chan;
    const int sc = a.s_begin + sl;
    float2 v = make_float2(0.f, 0.f);
    if (a.fj < __ldg(a.n_fac + sc)) {
        const int half = a.L / 2;
        const int kk = k <= half ? k : a.L - k;
        float2 x = a.transposed ? __ldg(a.X + (long long)c * (half + 1) + kk) : __ldg(a.X + (long long)kk * a.n_chan + c);
        if (k > half) x.y = -x.y;
        const float2 t = __ldg(a.kern + ((long long)sc * a.max_fac + a.fj) * a.L + k);
        v = cconj(cmul(x, t));
    }
    V[g] = v;
}

__device__ __forceinline__ float2 lcpow_real(float2 z, float ex) {
    const float r = hypotf(z.x, z.y);
    if (r == 0.f) return make_float2(0.f, 0.f);
    const float mag = powf(r, ex);
    float sn, cs;
    sincosf(ex * atan2f(z.y, z.x), &sn, &cs);
    return make_float2(mag * cs, mag * sn);
}

// y = conj(F[n][e]) for n < n_time: fold factor fj into the running product; the last factor converts and stores
__global__ void __launch_bounds__(LT) lcwt_fold_kernel(const CwtLongArgs a, const float2* __restrict__ F) {
    const int E = a.s_count * a.n_chan;
    const long long total = (long long)a.n_time * E;
    const long long g = (long long)blockIdx.x * LT + threadIdx.x;
    if (g >= total) return;
    const int n = (int)(g / E);
    const int e = (int)(g - (long long)n * E);
    const int sl = e / a.n_chan, c = e - sl * a.n_chan;
    const int sc = a.s_begin + sl;
    const int nf = __ldg(a.n_fac + sc);
    if (a.fj >= nf) return;
    const float ex = __ldg(a.expo + sc * a.max_fac + a.fj);
    const bool mag_only = nf > 1 && (a.out_kind == OUT_POW || a.out_kind == OUT_ABS);
    const float2 f = F[g];
    float2 z;
    if (mag_only) {
        const float l = 0.5f * ex * __log2f(f.x * f.x + f.y * f.y);
        z.x = a.fj == 0 ? l : a.zacc[g].x + l;
        z.y = 0.f;
    } else {
        float2 y = cconj(f);
        if (ex != 1.f) y = lcpow_real(y, ex);
        z = a.fj == 0 ? y : cmul(a.zacc[g], y);
    }
    if (a.fj < nf - 1) { a.zacc[g] = z; return; }
    const long long o = a.transposed ? ((long long)sc * a.n_chan + c) * a.n_time + n
                                     : ((long long)n * a.n_scales + sc) * a.n_chan + c;
    if (mag_only) {
        const float mag = exp2f(z.x);
        reinterpret_cast<float*>(a.out)[o] = a.out_kind == OUT_POW ? mag * mag : mag;
    } else if (a.out_kind == OUT_FOURIER) {
        reinterpret_cast<float2*>(a.out)[o] = z;
    } else {
        reinterpret_cast<float*>(a.out)[o] = convert_real(z, a.out_kind);
    }
}

}  // namespace

int cwt_factors_long(const CwtDesc& d, cudaStream_t st) {
    const int L = d.n_dft;
    if (L > MAX_LONG_LEN) return fail("cwt: circular length %d exceeds the supported maximum", L);
    if (L % 2) return fail("cwt: circular length must be even (got %d)", L);
    LongPlan plan = factorize(L);
    if (plan.max_prime > MAX_GENERIC_RADIX) return fail("cwt: circular length %d has a prime factor > %d", L, MAX_GENERIC_RADIX);
    Scratch sc(st);
    float2* tw = sc.take<float2>(L);
    if (!tw) return 1;
    ltwiddle_kernel<<<blocks_for(L), LT, 0, st>>>(tw, L);
    SPYB_LAUNCH_CHECK("ltwiddle_kernel");
    count_launch();
    long long sper = (long long)(WORK_BUDGET / ((size_t)L * d.n_chan * sizeof(float2)));
    if (sper < 1) sper = 1;
    if (sper > d.n_scales) sper = d.n_scales;
    const size_t ecap = (size_t)sper * d.n_chan;
    float2* bufA = sc.take<float2>((size_t)L * ecap);
    float2* bufB = sc.take<float2>((size_t)L * ecap);
    float2* zacc = d.max_fac > 1 ? sc.take<float2>((size_t)d.n_time * ecap) : nullptr;
    if (!bufA || !bufB || (d.max_fac > 1 && !zacc)) return 1;
    // the number of factors per scale lives on the device; every chunk runs max_fac rounds and the kernels skip
    // the scales that have fewer
    const size_t out_elem = d.out_kind == OUT_FOURIER ? 8 : 4;
    const int half1 = L / 2 + 1;
    for (int trial = 0; trial < d.n_trials; ++trial) {
        CwtLongArgs a;
        a.X = reinterpret_cast<const float2*>(d.xspec) + (long long)trial * half1 * d.n_chan;
        a.transposed = d.transposed;
        a.kern = reinterpret_cast<const float2*>(d.kern);
        a.expo = d.expo; a.n_fac = d.n_fac;
        a.n_chan = d.n_chan; a.L = L; a.max_fac = d.max_fac; a.n_scales = d.n_scales;
        a.n_time = d.n_time; a.out_kind = d.out_kind;
        a.zacc = zacc;
        a.out = static_cast<char*>(d.out) + (size_t)trial * d.n_time * d.n_scales * d.n_chan * out_elem;
        for (int s0 = 0; s0 < d.n_scales; s0 += (int)sper) {
            a.s_begin = s0;
            a.s_count = (int)(s0 + sper <= d.n_scales ? sper : d.n_scales - s0);
            const int E = a.s_count * d.n_chan;
            for (int fj = 0; fj < d.max_fac; ++fj) {
                a.fj = fj;
                lcwt_mul_kernel<<<blocks_for((long long)L * E), LT, 0, st>>>(a, bufA);
                SPYB_LAUNCH_CHECK("lcwt_mul_kernel");
                count_launch();
                float2* res = nullptr;
                if (lfft_axis0(bufA, bufB, tw, plan, E, &res, st)) return 1;
                lcwt_fold_kernel<<<blocks_for((long long)d.n_time * E), LT, 0, st>>>(a, res);
                SPYB_LAUNCH_CHECK("lcwt_fold_kernel");
                count_launch();
            }
        }
    }
    return 0;
}

}  // namespace spyb
