// K1 / K4: fused detrend -> taper -> real FFT -> scale -> frequency gather -> output
// conversion (-> taper mean) for whole trials (mtmfft) and for sliding frames (mtmconvol).
//
// Replaces the NumPy bodies of
//   syncopy/specest/mtmfft.py:96-127      (window, de-mean, rfft, _norm_spec)
//   syncopy/specest/stft.py:101-157       (boundary / end padding, framing, per-segment detrend,
//                                          window, rfft, _norm_spec)
//   syncopy/specest/compRoutines.py:169-189, 386-413  (detrend, gather, conversion, taper mean)
//
// Data layout: trials are time-major [trial][sample][channel] float32 (AnalogData default
// dimord).  Two adjacent channels (a "pair") are packed as one complex series
// z[n] = x[n][c] + i x[n][c+1] -- which is literally the float2 found in memory -- so one
// complex FFT of length nfft yields both real spectra:
//   X_c[k] = (Z[k] + conj Z[-k]) / 2,   X_{c+1}[k] = (Z[k] - conj Z[-k]) / (2i).
// A block handles P pairs (2P channels) of G frames; lanes are interleaved pair-fastest so a
// warp reads (32/P) consecutive time samples x (8P) contiguous bytes.
//
// Arbitrary (non power-of-two, odd) lengths run through Bluestein's chirp-z on the same block
// FFT: a[i] = z[i] conj(b[i]), Z[k] = conj(b[k]) * IFFT(FFT(a) * FFT(b))[k], b[i] = e^{i pi i^2/n}.
#include "common.cuh"
#include "fft_core.cuh"
#include "mtm_args.cuh"
#include "plan.cuh"
#include "spyb_internal.h"

#include <cstdlib>

namespace spyb {


// Sum `NV` values over the NT threads that share (g, p).  red: [nwarps][P][NV] floats.
template <int NT, int P, int NV>
__device__ __forceinline__ void group_reduce(float (&val)[NV], float* red, int tid, int p) {
    constexpr int GT = NT * P;                       // threads per group
    constexpr int SH = GT < 32 ? GT : 32;
#pragma unroll
    for (int off = P; off < SH; off <<= 1) {
#pragma unroll
        for (int i = 0; i < NV; ++i) val[i] += __shfl_xor_sync(0xffffffffu, val[i], off);
    }
    if constexpr (GT > 32) {
        constexpr int WG = GT / 32;                  // warps per group
        const int warp = tid >> 5, lane = tid & 31;
        __syncthreads();                             // red may still be read from a previous call
        if (lane < P) {
#pragma unroll
            for (int i = 0; i < NV; ++i) red[(warp * P + lane) * NV + i] = val[i];
        }
        __syncthreads();
        const int w0 = (warp / WG) * WG;
#pragma unroll
        for (int i = 0; i < NV; ++i) val[i] = 0.f;
        for (int w = 0; w < WG; ++w) {
#pragma unroll
            for (int i = 0; i < NV; ++i) val[i] += red[((w0 + w) * P + p) * NV + i];
        }
    }
}

template <int LOG2N, int P, bool BLUE, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) mtm_kernel(const MtmArgs a) {
    constexpr int N = 1 << LOG2N;
    constexpr int NT = N / 16;
    constexpr int SPAD = fft_padded_len(N);
    extern __shared__ float2 smem[];

    const int tid = threadIdx.x;
    const int p = tid % P;
    const int j = (tid / P) % NT;
    const int g = tid / (P * NT);
    const int G = blockDim.x / (P * NT);
    float2* s = smem + (size_t)g * SPAD * P;
    float* red = reinterpret_cast<float*>(smem + (size_t)G * SPAD * P);

    int frame = blockIdx.y * G + g;
    const bool frame_ok = frame < a.n_frames;
    if (!frame_ok) frame = a.n_frames - 1;           // keep the block converged; stores are masked
    const int trial = blockIdx.z;
    const int c = (blockIdx.x * P + p) * 2;
    const bool ca_ok = c < a.n_chan, cb_ok = c + 1 < a.n_chan;
    const long long start = (long long)a.frame_start0 + (long long)frame * a.hop;
    const float* __restrict__ xt = a.x + (long long)trial * a.trial_stride;
    const int n_win = a.n_win;

    float2 v[16];

    auto load_raw = [&]() {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int n = j + NT * e;
            const long long m = start + n;
            float2 val = make_float2(0.f, 0.f);
            if (n < n_win && m >= 0 && m < a.n_samples && ca_ok) {
                const float* ptr = xt + m * a.n_chan + c;
                if (a.vec_in) {
                    val = __ldg(reinterpret_cast<const float2*>(ptr));
                } else {
                    val.x = __ldg(ptr);
                    if (cb_ok) val.y = __ldg(ptr + 1);
                }
            }
            v[e] = val;
        }
    };

    // ---- detrending statistics over the window (scipy.signal.detrend, constant / linear) ----
    float mean_a = 0.f, mean_b = 0.f, slope_a = 0.f, slope_b = 0.f;
    const float tmid = 0.5f * (float)(n_win - 1);
    bool have_raw = false;
    if (a.polyremoval >= 0) {
        load_raw();
        have_raw = true;
        float sums[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int n = j + NT * e;
            if (n < n_win) {
                const float t = (float)n - tmid;
                sums[0] += v[e].x; sums[1] += v[e].y;
                sums[2] += t * v[e].x; sums[3] += t * v[e].y;
            }
        }
        group_reduce<NT, P, 4>(sums, red, tid, p);
        const float inv_n = 1.f / (float)n_win;
        mean_a = sums[0] * inv_n; mean_b = sums[1] * inv_n;
        if (a.polyremoval == 1 && n_win > 1) {
            // sum t^2 = n (n^2 - 1) / 12 for t centred
            const float stt = (float)((double)n_win * ((double)n_win * n_win - 1.0) / 12.0);
            slope_a = sums[2] / stt; slope_b = sums[3] / stt;
        }
    }

    const float half_scale = 0.5f * a.scale;
    const float inv_ntap = 1.f / (float)a.n_tapers;
    float amax_a = 0.f, amax_b = 0.f;

    for (int k = 0; k < a.n_tapers; ++k) {
        if (!have_raw) load_raw();
        have_raw = false;
        const float* __restrict__ win = a.tapers + (long long)k * n_win;

        // ---- detrend + taper ----
        float tsum[2] = {0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int n = j + NT * e;
            if (n < n_win) {
                const float t = (float)n - tmid;
                const float w = __ldg(win + n);
                v[e].x = (v[e].x - (mean_a + slope_a * t)) * w;
                v[e].y = (v[e].y - (mean_b + slope_b * t)) * w;
                tsum[0] += v[e].x; tsum[1] += v[e].y;
            }
        }
        if (a.demean_taper) {
            group_reduce<NT, P, 2>(tsum, red, tid, p);
            const float ma = tsum[0] / (float)n_win, mb = tsum[1] / (float)n_win;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n = j + NT * e;
                if (n < n_win) { v[e].x -= ma; v[e].y -= mb; }
            }
        }

        if constexpr (BLUE) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n = j + NT * e;
                if (n < n_win) v[e] = cmulc(v[e], __ldg(a.chirp + n));
            }
        }

        // ---- FFT (natural-order result lands in shared memory) ----
        block_fft<LOG2N, P>(v, s, a.tw, j, p);

        if constexpr (BLUE) {
            // Y = FFT(conj(A * bhat)); the caller-visible Z[k] = conj(b[k] * Y[k])
            fft_gather<N, P>(v, s, j, p);
            __syncthreads();
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = cconj(cmul(v[e], __ldg(a.bhat + j + NT * e)));
            block_fft<LOG2N, P>(v, s, a.tw, j, p);
        }

        // ---- epilogue: split the pair, scale, gather, convert, store ----
        const int n_dft = a.n_dft;
        for (int fi = j; fi < a.n_freq_out; fi += NT) {
            const int kf = a.freq_idx ? __ldg(a.freq_idx + fi) : fi;
            const int kn = kf == 0 ? 0 : n_dft - kf;
            float2 z1 = s[fft_pad(kf) * P + p];
            float2 z2 = s[fft_pad(kn) * P + p];
            if constexpr (BLUE) {
                z1 = cconj(cmul(z1, __ldg(a.chirp + kf)));
                z2 = cconj(cmul(z2, __ldg(a.chirp + kn)));
            }
            const float2 xa = make_float2((z1.x + z2.x) * half_scale, (z1.y - z2.y) * half_scale);
            const float2 xb = make_float2((z1.y + z2.y) * half_scale, (z2.x - z1.x) * half_scale);
            amax_a = fmaxf(amax_a, fmaxf(fabsf(xa.x), fabsf(xa.y)));
            amax_b = fmaxf(amax_b, fmaxf(fabsf(xb.x), fabsf(xb.y)));
            if (!frame_ok || !ca_ok) continue;
            const long long off = (long long)trial * a.so_trial + (long long)frame * a.so_frame +
                                  (a.keeptapers ? (long long)k * a.so_taper : 0LL) +
                                  (long long)fi * a.so_freq + c;
            const bool first = a.keeptapers || k == 0;
            const bool lastk = !a.keeptapers && k == a.n_tapers - 1 && a.n_tapers > 1;
            if (a.out_kind == OUT_FOURIER_PLANAR) {
                // strides are in floats; re plane at `off`, im plane n_chan floats later (keeptapers only)
                float* o = reinterpret_cast<float*>(a.out) + off;
                if (a.vec_out && cb_ok) {
                    *reinterpret_cast<float2*>(o) = make_float2(xa.x, xb.x);
                    *reinterpret_cast<float2*>(o + a.n_chan) = make_float2(xa.y, xb.y);
                } else {
                    o[0] = xa.x; o[a.n_chan] = xa.y;
                    if (cb_ok) { o[1] = xb.x; o[a.n_chan + 1] = xb.y; }
                }
            } else if (a.out_kind == OUT_FOURIER) {
                float2* o = reinterpret_cast<float2*>(a.out) + off;
                float2 ra = xa, rb = xb;
                if (a.vec_out && cb_ok) {
                    float4* o4 = reinterpret_cast<float4*>(o);
                    if (!first) { const float4 old = *o4; ra.x += old.x; ra.y += old.y; rb.x += old.z; rb.y += old.w; }
                    if (lastk) { ra.x *= inv_ntap; ra.y *= inv_ntap; rb.x *= inv_ntap; rb.y *= inv_ntap; }
                    *o4 = make_float4(ra.x, ra.y, rb.x, rb.y);
                } else {
                    if (!first) { const float2 old = o[0]; ra.x += old.x; ra.y += old.y; }
                    if (lastk) { ra.x *= inv_ntap; ra.y *= inv_ntap; }
                    o[0] = ra;
                    if (cb_ok) {
                        if (!first) { const float2 old = o[1]; rb.x += old.x; rb.y += old.y; }
                        if (lastk) { rb.x *= inv_ntap; rb.y *= inv_ntap; }
                        o[1] = rb;
                    }
                }
            } else {
                float* o = reinterpret_cast<float*>(a.out) + off;
                float ra = convert_real(xa, a.out_kind), rb = convert_real(xb, a.out_kind);
                if (a.vec_out && cb_ok) {
                    float2* o2 = reinterpret_cast<float2*>(o);
                    if (!first) { const float2 old = *o2; ra += old.x; rb += old.y; }
                    if (lastk) { ra *= inv_ntap; rb *= inv_ntap; }
                    *o2 = make_float2(ra, rb);
                } else {
                    if (!first) ra += o[0];
                    if (lastk) ra *= inv_ntap;
                    o[0] = ra;
                    if (cb_ok) {
                        if (!first) rb += o[1];
                        if (lastk) rb *= inv_ntap;
                        o[1] = rb;
                    }
                }
            }
        }
        __syncthreads();   // the next taper (or the Bluestein pass) overwrites the exchange buffer
    }

    if (a.chan_amax != nullptr) {
        float am[2] = {amax_a, amax_b};
        // max-reduce over the group: reuse the shuffle pattern with fmaxf
        constexpr int GT = NT * P;
        constexpr int SH = GT < 32 ? GT : 32;
#pragma unroll
        for (int off = P; off < SH; off <<= 1) {
            am[0] = fmaxf(am[0], __shfl_xor_sync(0xffffffffu, am[0], off));
            am[1] = fmaxf(am[1], __shfl_xor_sync(0xffffffffu, am[1], off));
        }
        // non-negative floats order like their bit patterns
        if (((tid & 31) < P) && frame_ok) {
            if (ca_ok) atomicMax(reinterpret_cast<int*>(a.chan_amax) + c, __float_as_int(am[0]));
            if (cb_ok) atomicMax(reinterpret_cast<int*>(a.chan_amax) + c + 1, __float_as_int(am[1]));
        }
    }
}

// ---------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------
template <int LOG2N, int P, bool BLUE>
static int launch_one(const MtmArgs& a, cudaStream_t stream) {
    constexpr int N = 1 << LOG2N, NT = N / 16, GT = NT * P;
    constexpr int G = GT >= 256 ? 1 : 256 / GT;
    constexpr int THREADS = GT * G;
    // 16 complex values per thread stay in registers across the passes: leave room for ~100 registers per thread
    // wherever the thread count allows it (a 64-register cap spills them to local memory)
    constexpr int MINB = THREADS >= 512 ? 1 : 512 / THREADS;
    auto kern = mtm_kernel<LOG2N, P, BLUE, THREADS, MINB>;
    const size_t smem = (size_t)G * fft_padded_len(N) * P * sizeof(float2) +
                        (size_t)(THREADS / 32 + 1) * P * 4 * sizeof(float);
    // per device / context attribute: set on every launch (several engines may live in one process)
    SPYB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int chan_tiles = (a.n_chan + 2 * P - 1) / (2 * P);
    const int frame_blocks = (a.n_frames + G - 1) / G;
    if (frame_blocks > 65535 || a.n_trials > 65535)
        return fail("mtm launch: too many frames (%d) or trials (%d) for one launch", a.n_frames, a.n_trials);
    dim3 grid(chan_tiles, frame_blocks, a.n_trials);
    kern<<<grid, THREADS, smem, stream>>>(a);
    SPYB_LAUNCH_CHECK("mtm_kernel");
    count_launch();
    return 0;
}

template <bool BLUE>
static int launch_log2(int log2n, const MtmArgs& a, cudaStream_t st) {
    switch (log2n) {
        case 4:  return launch_one<4, 4, BLUE>(a, st);
        case 5:  return launch_one<5, 4, BLUE>(a, st);
        case 6:  return launch_one<6, 4, BLUE>(a, st);
        case 7:  return launch_one<7, 4, BLUE>(a, st);
        case 8:  return launch_one<8, 4, BLUE>(a, st);
        case 9:  return launch_one<9, 4, BLUE>(a, st);
        case 10: return launch_one<10, 4, BLUE>(a, st);
        case 11: return launch_one<11, 4, BLUE>(a, st);
        case 12: return launch_one<12, 2, BLUE>(a, st);     // 512 threads: room for ~100 registers, no spills
        case 13: return launch_one<13, 2, BLUE>(a, st);
        case 14: return launch_one<14, 1, BLUE>(a, st);
        default: return fail("unsupported block FFT size 2^%d", log2n);
    }
}

int mtm_frames(const MtmFramesDesc& d, cudaStream_t stream) {
    if (d.n_trials <= 0 || d.n_frames <= 0 || d.n_chan <= 0 || d.n_freq_out <= 0) return 0;
    if (d.n_win < 1 || d.n_win > d.n_dft) return fail("window length %d must be in [1, nfft=%d]", d.n_win, d.n_dft);
    if (d.n_tapers < 1) return fail("need at least one taper");
    if (mtm_needs_long(d.n_dft)) return mtm_frames_long(d, stream);     // beyond the shared-memory kernels
    const FftPlan* pl = get_fft_plan(d.n_dft);
    if (!pl) return 1;

    MtmArgs a;
    a.x = d.x; a.trial_stride = d.trial_stride;
    a.n_trials = d.n_trials; a.n_samples = d.n_samples; a.n_chan = d.n_chan;
    a.n_win = d.n_win; a.n_dft = d.n_dft;
    a.frame_start0 = d.frame_start0; a.hop = d.hop; a.n_frames = d.n_frames;
    a.tapers = d.tapers; a.n_tapers = d.n_tapers;
    a.polyremoval = d.polyremoval; a.demean_taper = d.demean_taper; a.scale = d.scale;
    a.freq_idx = d.freq_idx; a.n_freq_out = d.n_freq_out;
    a.out_kind = d.out_kind; a.keeptapers = d.keeptapers;
    a.out = d.out;
    a.so_trial = d.so_trial; a.so_frame = d.so_frame; a.so_taper = d.so_taper; a.so_freq = d.so_freq;
    a.chan_amax = d.chan_amax;
    a.tw = pl->tw; a.chirp = pl->chirp; a.bhat = pl->bhat;

    const bool even_in = (d.n_chan % 2 == 0) && (d.trial_stride % 2 == 0) &&
                         (reinterpret_cast<uintptr_t>(d.x) % 8 == 0);
    a.vec_in = even_in ? 1 : 0;
    if (d.out_kind == OUT_FOURIER_PLANAR && !d.keeptapers)
        return fail("planar complex output needs keeptapers = 1");
    const size_t elem = out_is_complex(d.out_kind) ? 8 : 4;
    const bool even_out = (d.so_trial % 2 == 0) && (d.so_frame % 2 == 0) && (d.so_taper % 2 == 0) &&
                          (d.so_freq % 2 == 0) && (reinterpret_cast<uintptr_t>(d.out) % (2 * elem) == 0);
    a.vec_out = even_out ? 1 : 0;
    a.vec16 = ((d.n_chan % 4 == 0) && (d.trial_stride % 4 == 0) && (reinterpret_cast<uintptr_t>(d.x) % 16 == 0)) ? 1 : 0;
    a.tw_dif = pl->tw_dif;

    // power-of-two lengths >= 256 run on the shared-memory-resident in-place kernel (mtm_dif.cu); short
    // windows (several frames per block) and Bluestein lengths on the Stockham kernel below
    static const bool force_stockham = getenv("SPYB_MTM_STOCKHAM") != nullptr;
    if (!pl->bluestein && !force_stockham) {
        const int r4 = mtm_launch_4s(pl->log2n, a, stream);
        if (r4 >= 0) return r4;
        const int rt = mtm_launch_tma(pl->log2n, a, stream);
        if (rt >= 0) return rt;
        const int r8 = mtm_launch_r8(pl->log2n, a, stream);
        if (r8 >= 0) return r8;
        const int rc = mtm_launch_dif(pl->log2n, a, stream);
        if (rc >= 0) return rc;
    }
    return pl->bluestein ? launch_log2<true>(pl->log2n, a, stream) : launch_log2<false>(pl->log2n, a, stream);
}

}  // namespace spyb
