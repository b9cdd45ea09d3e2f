// K1 / K4 for power-of-two lengths: the tile stays in shared memory, the FFT runs in place.
//
// Same arithmetic as mtm.cu (detrend -> taper -> [de-mean] -> real FFT of two channels packed as one complex
// series -> scale -> gather -> convert -> [taper mean]); different execution shape, chosen for HBM throughput:
//
//   * a block owns P channel pairs of one frame: N x 2P floats = N*P*8 bytes of shared memory, filled with 16-byte
//     row copies (the detrending sums ride along), so for N = 4096 several 64 KB blocks share an SM and their
//     load / FFT / store phases overlap;
//   * radix-16 decimation-in-frequency passes work in place (gather 16 values, butterfly, twiddle, write back to
//     the same 16 slots), so a thread can process several butterflies of a pass one after the other with only 16
//     complex values live -- the thread count is decoupled from N and ~80 registers suffice (the auto-sorting
//     Stockham kernel needs all values of a pass in registers at once and ran into a 64-register spill wall);
//   * the spectrum comes out digit-reversed; the epilogue reads bin k at dif_pos(k);
//   * shared-memory index = swz(i) * P + p with swz(i) = i ^ nibble1(i) ^ nibble2(i) ^ nibble3(i): every access
//     pattern of the kernel (row copies, the three pass strides, the digit-reversed epilogue) touches 16
//     distinct slots per 16 consecutive logical indices, i.e. no bank conflicts beyond the 2-wavefront minimum.
#include "common.cuh"
#include "fft_core.cuh"
#include "mtm_args.cuh"
#include "spyb_internal.h"

#include <cstdlib>

namespace spyb {
namespace {

__host__ __device__ __forceinline__ constexpr int swz(int i) {
    return i ^ ((i >> 4) & 15) ^ ((i >> 8) & 15) ^ ((i >> 12) & 15);
}

// slot of frequency bin k after the passes (radices 16, ..., 16, 2^(LOG2N % 4))
template <int LOG2N>
__device__ __forceinline__ int dif_pos(int k) {
    constexpr int Q16 = LOG2N / 4;
    int pos = 0, sh = LOG2N;
#pragma unroll
    for (int i = 0; i < Q16; ++i) {
        sh -= 4;
        pos |= (k & 15) << sh;
        k >>= 4;
    }
    return pos | k;
}

// One in-place DIF pass of radix R over butterflies at distance STRIDE.
// swz() is linear over XOR and base has zero bits where r*STRIDE lives, so the slot of element r is
// swz(base) ^ swz(r*STRIDE): one XOR with a compile-time constant per shared-memory access.
template <int LOG2N, int P, int THREADS, int R, int STRIDE, bool FIRST, typename Pre>
__device__ __forceinline__ void dif_pass(float2* __restrict__ s, const float2* __restrict__ tw, int tid, Pre& pre) {
    constexpr int N = 1 << LOG2N;
    constexpr int ITEMS = (N / R) * P;
    char* sb = reinterpret_cast<char*>(s);
#pragma unroll 1
    for (int item = tid; item < ITEMS; item += THREADS) {
        const unsigned p = (unsigned)item % P, u = (unsigned)item / P;
        const unsigned o = u % STRIDE, base = (u / STRIDE) * (R * STRIDE) + o;
        const unsigned a0 = ((unsigned)swz((int)base) * P + p) * 8u;
        float2 x[R];
#pragma unroll
        for (int r = 0; r < R; ++r) x[r] = *reinterpret_cast<const float2*>(sb + (a0 ^ (unsigned)(swz(r * STRIDE) * P * 8)));
        if constexpr (FIRST) pre(x, (int)base);
        Radix<R>::run(x);
        if constexpr (STRIDE > 1) {
            const float2* __restrict__ two = tw + o;
#pragma unroll
            for (int q = 1; q < R; ++q) {
                const int reg = Radix<R>::reg_of(q);
                x[reg] = cmul(x[reg], __ldg(two + (q - 1) * STRIDE));
            }
        }
#pragma unroll
        for (int q = 0; q < R; ++q)
            *reinterpret_cast<float2*>(sb + (a0 ^ (unsigned)(swz(q * STRIDE) * P * 8))) = x[Radix<R>::reg_of(q)];
    }
    __syncthreads();
}

template <int LOG2N, int P, int THREADS, int I>
struct DifPasses {
    static constexpr int Q16 = LOG2N / 4;
    static constexpr int RL = 1 << (LOG2N % 4);
    template <typename Pre>
    __device__ __forceinline__ static void run(float2* s, const float2* tw, int tid, Pre& pre) {
        if constexpr (I < Q16) {
            constexpr int STRIDE = 1 << (LOG2N - 4 * (I + 1));
            dif_pass<LOG2N, P, THREADS, 16, STRIDE, I == 0>(s, tw, tid, pre);
            DifPasses<LOG2N, P, THREADS, I + 1>::run(s, tw + (STRIDE > 1 ? 15 * STRIDE : 0), tid, pre);
        } else if constexpr (RL > 1) {
            dif_pass<LOG2N, P, THREADS, RL, 1, false>(s, tw, tid, pre);
        }
    }
};

// Sum NV values over all threads of the block that share p = tid % P.  red: [THREADS/32][P][NV] floats.
template <int P, int THREADS, int NV>
__device__ __forceinline__ void pair_reduce(float (&val)[NV], float* red, int tid) {
    constexpr int SH = THREADS < 32 ? THREADS : 32;
#pragma unroll
    for (int off = P; off < SH; off <<= 1) {
#pragma unroll
        for (int i = 0; i < NV; ++i) val[i] += __shfl_xor_sync(0xffffffffu, val[i], off);
    }
    if constexpr (THREADS > 32) {
        constexpr int NW = THREADS / 32;
        const int warp = tid >> 5, lane = tid & 31, p = tid % P;
        __syncthreads();
        if (lane < P) {
#pragma unroll
            for (int i = 0; i < NV; ++i) red[(warp * P + lane) * NV + i] = val[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < NV; ++i) val[i] = 0.f;
        for (int w = 0; w < NW; ++w) {
#pragma unroll
            for (int i = 0; i < NV; ++i) val[i] += red[(w * P + p) * NV + i];
        }
    }
}

template <int LOG2N, int P, int THREADS, int MINB, bool ACC>
__global__ void __launch_bounds__(THREADS, MINB) mtm_dif_kernel(const MtmArgs a) {
    constexpr int N = 1 << LOG2N;
    constexpr int STRIDE0 = N / 16;
    static_assert(THREADS % P == 0 && THREADS % 32 == 0, "a thread must keep its pair across loop iterations");
    extern __shared__ __align__(16) float2 smem[];
    float2* s = smem;
    float* red = reinterpret_cast<float*>(smem + (size_t)N * P);

    const int tid = threadIdx.x;
    const int p = tid % P;
    const int frame = blockIdx.y, trial = blockIdx.z;
    const int c0 = blockIdx.x * 2 * P;
    const int c = c0 + 2 * p;
    const bool ca_ok = c < a.n_chan, cb_ok = c + 1 < a.n_chan;
    const bool full_tile = c0 + 2 * P <= a.n_chan;
    const long long start = (long long)a.frame_start0 + (long long)frame * a.hop;
    const float* __restrict__ xt = a.x + (long long)trial * a.trial_stride;
    const int n_win = a.n_win;

    // ---- raw tile -> shared memory (zeros outside the window / the trial) ----
    // 16-byte rows go through registers (LDG.128 -> STS.128: 4 shared-memory wavefronts per warp; cp.async scatters
    // one wavefront per lane here because every lane's 16 bytes come from a different 1 KB-strided row), which
    // also lets the detrending sums ride along with the first copy.
    const float tmid = 0.5f * (float)(n_win - 1);
    const bool vec_tile = P >= 2 && a.vec16 && full_tile;
    constexpr int CH = P >= 2 ? P / 2 : 1;                          // 16-byte chunks per row
    constexpr int RED_PER_WARP = CH * 8 > P * 4 ? CH * 8 : P * 4;
    float* accb = red + (THREADS / 32) * RED_PER_WARP + 4 * P;      // taper-mean accumulators (acc_smem)
    auto load_tile = [&](float (&sums)[8], const bool with_sums) {
        if (vec_tile) {
            const int h = tid % CH;
            constexpr int ROWS_PER_STEP = THREADS / CH;
            // every 16-byte load of the thread is issued before the first store: one HBM latency per tile
            constexpr int UN = (N / ROWS_PER_STEP) < 16 ? (N / ROWS_PER_STEP) : 16;
            const float* __restrict__ src = xt + c0 + 4 * h;
            for (int n0 = tid / CH; n0 < N; n0 += UN * ROWS_PER_STEP) {
                float4 v[UN];
#pragma unroll
                for (int i = 0; i < UN; ++i) {
                    const int n = n0 + i * ROWS_PER_STEP;
                    const long long m = start + n;
                    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (n < n_win && m >= 0 && m < a.n_samples)
                        v[i] = __ldg(reinterpret_cast<const float4*>(src + m * a.n_chan));
                }
#pragma unroll
                for (int i = 0; i < UN; ++i) {
                    const int n = n0 + i * ROWS_PER_STEP;
                    if (n < N) *reinterpret_cast<float4*>(&s[swz(n) * P + 2 * h]) = v[i];
                    if (with_sums && n < n_win) {
                        const float t = (float)n - tmid;
                        sums[0] += v[i].x; sums[1] += v[i].y; sums[2] += v[i].z; sums[3] += v[i].w;
                        sums[4] += t * v[i].x; sums[5] += t * v[i].y; sums[6] += t * v[i].z; sums[7] += t * v[i].w;
                    }
                }
            }
        } else {
#pragma unroll 4
            for (int q = tid; q < N * P; q += THREADS) {
                const int n = q / P;                                // q % P == p
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
                s[swz(n) * P + p] = val;
                if (with_sums && n < n_win) {
                    const float t = (float)n - tmid;
                    sums[0] += val.x; sums[1] += val.y;
                    sums[4] += t * val.x; sums[5] += t * val.y;
                }
            }
        }
        __syncthreads();
    };

    // ---- detrending statistics over the window (scipy.signal.detrend, constant / linear) ----
    float mean_a = 0.f, mean_b = 0.f, slope_a = 0.f, slope_b = 0.f;
    {
        float sums[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const bool want = a.polyremoval >= 0;
        load_tile(sums, want);
        if (want) {
            float* stat = red + (THREADS / 32) * RED_PER_WARP;      // [2P channels][2]
            if (vec_tile) {
                // sums are per 16-byte chunk (4 channels); reduce over the threads that share the chunk
                pair_reduce<CH, THREADS, 8>(sums, red, tid);
                if (tid < CH) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) { stat[(4 * tid + i) * 2] = sums[i]; stat[(4 * tid + i) * 2 + 1] = sums[4 + i]; }
                }
            } else {
                float s4[4] = {sums[0], sums[1], sums[4], sums[5]};
                pair_reduce<P, THREADS, 4>(s4, red, tid);
                if (tid < P) {
                    stat[(2 * tid) * 2] = s4[0]; stat[(2 * tid + 1) * 2] = s4[1];
                    stat[(2 * tid) * 2 + 1] = s4[2]; stat[(2 * tid + 1) * 2 + 1] = s4[3];
                }
            }
            __syncthreads();
            const float inv_n = 1.f / (float)n_win;
            mean_a = stat[(2 * p) * 2] * inv_n; mean_b = stat[(2 * p + 1) * 2] * inv_n;
            if (a.polyremoval == 1 && n_win > 1) {
                const float stt = (float)((double)n_win * ((double)n_win * n_win - 1.0) / 12.0);
                slope_a = stat[(2 * p) * 2 + 1] / stt; slope_b = stat[(2 * p + 1) * 2 + 1] / stt;
            }
        }
    }

    const float half_scale = 0.5f * a.scale;
    const float inv_ntap = 1.f / (float)a.n_tapers;
    float amax_a = 0.f, amax_b = 0.f;

    for (int k = 0; k < a.n_tapers; ++k) {
        if (k > 0) {                                                // the passes overwrote the raw samples
            float dummy[8];
            load_tile(dummy, false);
        }
        const float* __restrict__ win = a.tapers + (long long)k * n_win;

        // mean of the tapered window (mtmfft.py:114-116), subtracted inside the first pass
        float tm_a = 0.f, tm_b = 0.f;
        if (a.demean_taper) {
            float tsum[2] = {0.f, 0.f};
            for (int q = tid; q < n_win * P; q += THREADS) {
                const int n = q / P;
                const float2 v = s[swz(n) * P + p];
                const float t = (float)n - tmid;
                const float w = __ldg(win + n);
                tsum[0] += (v.x - (mean_a + slope_a * t)) * w;
                tsum[1] += (v.y - (mean_b + slope_b * t)) * w;
            }
            pair_reduce<P, THREADS, 2>(tsum, red, tid);
            tm_a = tsum[0] / (float)n_win; tm_b = tsum[1] / (float)n_win;
        }

        const bool sloped = slope_a != 0.f || slope_b != 0.f;
        auto pre = [&](float2 (&x)[16], int base) {
            const float* __restrict__ wb = win + base;
            if (base + 15 * STRIDE0 < n_win && !sloped) {            // whole butterfly inside the window, no trend
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const float w = __ldg(wb + r * STRIDE0);
                    x[r].x = (x[r].x - mean_a) * w - tm_a;
                    x[r].y = (x[r].y - mean_b) * w - tm_b;
                }
            } else {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const int n = base + r * STRIDE0;
                    if (n < n_win) {
                        const float t = (float)n - tmid;
                        const float w = __ldg(wb + r * STRIDE0);
                        x[r].x = (x[r].x - (mean_a + slope_a * t)) * w - tm_a;
                        x[r].y = (x[r].y - (mean_b + slope_b * t)) * w - tm_b;
                    }
                }
            }
        };
        DifPasses<LOG2N, P, THREADS, 0>::run(s, a.tw_dif, tid, pre);

        // ---- epilogue: split the pair, scale, gather, convert, store ----
        const bool first = a.keeptapers || k == 0;
        const bool lastk = !a.keeptapers && k == a.n_tapers - 1 && a.n_tapers > 1;
        const long long off0 = (long long)trial * a.so_trial + (long long)frame * a.so_frame +
                               (a.keeptapers ? (long long)k * a.so_taper : 0LL) + c;
        for (int item = tid; item < a.n_freq_out * P; item += THREADS) {
            const int fi = item / P;
            const int kf = a.freq_idx ? __ldg(a.freq_idx + fi) : fi;
            const int kn = (N - kf) & (N - 1);
            const float2 z1 = s[swz(dif_pos<LOG2N>(kf)) * P + p];
            const float2 z2 = s[swz(dif_pos<LOG2N>(kn)) * P + p];
            const float2 xa = make_float2((z1.x + z2.x) * half_scale, (z1.y - z2.y) * half_scale);
            const float2 xb = make_float2((z1.y + z2.y) * half_scale, (z2.x - z1.x) * half_scale);
            amax_a = fmaxf(amax_a, fmaxf(fabsf(xa.x), fabsf(xa.y)));
            amax_b = fmaxf(amax_b, fmaxf(fabsf(xb.x), fabsf(xb.y)));
            if constexpr (ACC) {
                // taper mean without read-modify-write of the result: every thread revisits the same items for
                // every taper, so its running sums live in its own shared-memory slots
                if (a.out_kind == OUT_FOURIER) {
                    float4* ap = reinterpret_cast<float4*>(accb) + item;
                    float4 v = make_float4(xa.x, xa.y, xb.x, xb.y);
                    if (k > 0) { const float4 old = *ap; v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w; }
                    if (k < a.n_tapers - 1) { *ap = v; continue; }
                    if (!ca_ok) continue;
                    float2* o = reinterpret_cast<float2*>(a.out) + off0 + (long long)fi * a.so_freq;
                    o[0] = make_float2(v.x * inv_ntap, v.y * inv_ntap);
                    if (cb_ok) o[1] = make_float2(v.z * inv_ntap, v.w * inv_ntap);
                } else {
                    float2* ap = reinterpret_cast<float2*>(accb) + item;
                    float2 v = make_float2(convert_real(xa, a.out_kind), convert_real(xb, a.out_kind));
                    if (k > 0) { const float2 old = *ap; v.x += old.x; v.y += old.y; }
                    if (k < a.n_tapers - 1) { *ap = v; continue; }
                    if (!ca_ok) continue;
                    float* o = reinterpret_cast<float*>(a.out) + off0 + (long long)fi * a.so_freq;
                    if (a.vec_out && cb_ok) {
                        *reinterpret_cast<float2*>(o) = make_float2(v.x * inv_ntap, v.y * inv_ntap);
                    } else {
                        o[0] = v.x * inv_ntap;
                        if (cb_ok) o[1] = v.y * inv_ntap;
                    }
                }
                continue;
            }
            if (!ca_ok) continue;
            const long long off = off0 + (long long)fi * a.so_freq;
            if (a.out_kind == OUT_FOURIER_PLANAR) {
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
        __syncthreads();   // the next taper reloads the tile
    }

    if (a.chan_amax != nullptr) {
        constexpr int SH = THREADS < 32 ? THREADS : 32;
#pragma unroll
        for (int off = P; off < SH; off <<= 1) {
            amax_a = fmaxf(amax_a, __shfl_xor_sync(0xffffffffu, amax_a, off));
            amax_b = fmaxf(amax_b, __shfl_xor_sync(0xffffffffu, amax_b, off));
        }
        if ((tid & 31) < P) {      // non-negative floats order like their bit patterns
            if (ca_ok) atomicMax(reinterpret_cast<int*>(a.chan_amax) + c, __float_as_int(amax_a));
            if (cb_ok) atomicMax(reinterpret_cast<int*>(a.chan_amax) + c + 1, __float_as_int(amax_b));
        }
    }
}

template <int LOG2N, int P, int THREADS, int MINB>
int launch_dif(const MtmArgs& a_in, cudaStream_t stream) {
    constexpr int N = 1 << LOG2N;
    constexpr int CH = P >= 2 ? P / 2 : 1;
    constexpr int RED_PER_WARP = CH * 8 > P * 4 ? CH * 8 : P * 4;
    constexpr size_t kMaxSmem = 227 * 1024;
    size_t smem = (size_t)N * P * sizeof(float2) + (size_t)(THREADS / 32) * RED_PER_WARP * sizeof(float) +
                  (size_t)4 * P * sizeof(float) + 16;
    MtmArgs a = a_in;
    a.acc_smem = 0;
    if (!a.keeptapers && a.n_tapers > 1 && a.out_kind != OUT_FOURIER_PLANAR) {
        const size_t acc = (size_t)a.n_freq_out * P * (a.out_kind == OUT_FOURIER ? 16 : 8) + 16;
        if (smem + acc <= kMaxSmem) {
            smem = ((smem + 15) & ~(size_t)15) + acc;
            a.acc_smem = 1;
        }
    }
    auto kern = a.acc_smem ? mtm_dif_kernel<LOG2N, P, THREADS, MINB, true> : mtm_dif_kernel<LOG2N, P, THREADS, MINB, false>;
    // per device / context attribute: set on every launch (several engines may live in one process)
    SPYB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
    const int chan_tiles = (a.n_chan + 2 * P - 1) / (2 * P);
    if (a.n_frames > 65535 || a.n_trials > 65535)
        return fail("mtm launch: too many frames (%d) or trials (%d) for one launch", a.n_frames, a.n_trials);
    dim3 grid(chan_tiles, a.n_frames, a.n_trials);
    kern<<<grid, THREADS, smem, stream>>>(a);
    SPYB_LAUNCH_CHECK("mtm_dif_kernel");
    count_launch();
    return 0;
}

}  // namespace

int mtm_launch_dif(int log2n, const MtmArgs& a, cudaStream_t st) {
    // N = 4096: 4 pairs (32-byte rows) x 512 threads, one 128 KB block per SM measured fastest (0.78 ms at cfg-2);
    // 2 pairs x 256 threads x 3 blocks: 0.89 ms (half-sector row loads), 2 pairs x 256 x 2: 0.94 ms
    switch (log2n) {
        case 8:  return launch_dif<8, 4, 64, 12>(a, st);
        case 9:  return launch_dif<9, 4, 128, 6>(a, st);
        case 10: return launch_dif<10, 4, 256, 3>(a, st);
        case 11: return launch_dif<11, 4, 256, 3>(a, st);
        case 12: return launch_dif<12, 4, 512, 1>(a, st);
        case 13: return launch_dif<13, 2, 512, 1>(a, st);
        case 14: return launch_dif<14, 1, 512, 1>(a, st);
        default: return -1;
    }
}

}  // namespace spyb
