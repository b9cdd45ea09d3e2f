// K1 / K4 for 512-point windows (the sliding-window multitaper transform of BASELINE cfg-3: nperseg 512, 7 tapers):
// three radix-8 DIF passes, the raw frame stays in registers across all tapers.
//
// Same arithmetic as mtm_dif.cu (detrend -> taper -> [de-mean] -> real FFT of two channels packed as one complex
// series -> scale -> gather -> convert -> taper mean; syncopy/specest/stft.py:119-154, mtmconvol.py:136-150,
// compRoutines.py:410-413), built for the case where the FFT arithmetic, not HBM, is the bound: with K tapers every
// frame is transformed K times (cfg-3: 66 GFLOP against 1.7 GB per 100 trials).
//
//   * a tile = 8 channels (4 complex pairs) x 512 samples of one frame, owned by 128 threads; thread (o, h) holds rows
//     o + 64 r (r = 0..7) of the 16-byte column h in registers for the whole tile: the frame is read from global
//     memory ONCE, detrending sums come from the registers, and every taper starts from them (mtm_dif.cu re-reads
//     the tile per taper);
//   * 8 = 2^3 points per butterfly and two pairs per thread: complex adds / multiplies are packed FADD2 / FMUL2 /
//     FFMA2 (packed.cuh), shared memory is touched with 128-bit accesses only, the first pass needs no shared-memory
//     read at all;
//   * slot index = i ^ ((i >> 3) & 3): a quarter warp (8 lanes = 4 butterflies x 2 columns) covers a 128-byte line in
//     the second and third pass and in the digit-reversed epilogue reads;
//   * the twiddle tables (W_512^(o q), W_64^(o' q), q = 1..7) are computed once per block and sit in shared memory
//     next to the window table of all tapers;
//   * with keeptapers = 0 the running taper sums of a thread's output bins live in registers: one store per bin.
// A block works on 2 tiles (256 threads, two blocks per SM) that share the window table.
#include <cuda_runtime.h>

#include <cstdlib>

#include "common.cuh"
#include "mtm_args.cuh"
#include "packed.cuh"
#include "spyb_internal.h"

namespace spyb {
namespace {

constexpr int N = 512;
constexpr int TILE_THREADS = 128;          // N / 8 butterflies x 2 columns
constexpr int TILES = 2;                   // tiles per block
constexpr int THREADS = TILE_THREADS * TILES;
constexpr int ROW_BYTES = 32;
constexpr int TILE_BYTES = N * ROW_BYTES;  // 16 KB
constexpr int MAX_TAPERS_SMEM = 16;        // window table [K][512] floats in shared memory

__host__ __device__ constexpr int swz(int i) { return i ^ ((i >> 3) & 3); }

struct R8Args {
    MtmArgs m;
    int chan_tiles;
    long long n_tiles;
};

// KIND: 0 real-valued result (pow, abs, ...), 1 interleaved complex, 2 planar complex (out_kind 8)
template <int KIND>
__global__ void __launch_bounds__(THREADS, 2) mtm_r8_kernel(const R8Args ra) {
    extern __shared__ __align__(128) unsigned char smem[];
    const MtmArgs& a = ra.m;
    const int tid = threadIdx.x;
    const int tl = tid / TILE_THREADS, tt = tid % TILE_THREADS;      // tile of the block, thread of the tile
    const int h = tt & 1, o = tt >> 1;                               // 16-byte column, butterfly 0..63
    const int lane = tid & 31, wt = tt >> 5;                         // warp of the tile 0..3
    unsigned char* work = smem + tl * TILE_BYTES;
    float* wins = reinterpret_cast<float*>(smem + TILES * TILE_BYTES);                     // [K][512]
    float* red = wins + a.n_tapers * N + tl * (4 * 2 * 8);                                  // per tile [4 warps][2][8]
    float2* tw1 = reinterpret_cast<float2*>(wins + a.n_tapers * N + TILES * (4 * 2 * 8));   // [7][64]  W_512^(o q)
    float2* tw2 = tw1 + 7 * 64;                                                             // [7][8]   W_64^(o' q)
    const int K = a.n_tapers, n_win = a.n_win;

    // window table (zero past the window), shared by the block's tiles
    for (int i = tid; i < K * N; i += THREADS) {
        const int k = i / N, n = i % N;
        wins[i] = n < n_win ? __ldg(a.tapers + (long long)k * n_win + n) : 0.f;
    }
    // twiddle tables: pass 1 W_512^(o q), pass 2 W_64^(o' q), q = 1..7
    for (int i = tid; i < 7 * 64 + 7 * 8; i += THREADS) {
        float sn, cs;
        if (i < 7 * 64) {
            const int q = i / 64 + 1, oo = i % 64;
            sincospif(-2.0f * (float)((oo * q) % N) / (float)N, &sn, &cs);
        } else {
            const int q = (i - 7 * 64) / 8 + 1, oo = (i - 7 * 64) % 8;
            sincospif(-2.0f * (float)((oo * q) % 64) / 64.0f, &sn, &cs);
        }
        tw1[i] = make_float2(cs, sn);
    }
    __syncthreads();

    const long long tile = (long long)blockIdx.x * TILES + tl;
    const bool active = tile < ra.n_tiles;
    const long long tclamp = active ? tile : ra.n_tiles - 1;          // idle tiles redo the last one (barriers stay uniform)
    const int ct = (int)(tclamp % ra.chan_tiles);
    const long long tf = tclamp / ra.chan_tiles;
    const int frame = (int)(tf % a.n_frames), trial = (int)(tf / a.n_frames);
    const int c0 = ct * 8;
    const long long start = (long long)a.frame_start0 + (long long)frame * a.hop;

    // ---- the frame: 8 rows per thread, zeros outside the window / the trial ----
    c2 xa[8], xb[8];
    {
        const float* __restrict__ src = a.x + (long long)trial * a.trial_stride + c0 + 4 * h;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int n = o + 64 * r;
            const long long m = start + n;
            ulonglong2 v = make_ulonglong2(0ull, 0ull);
            if (n < n_win && m >= 0 && m < a.n_samples) v = *reinterpret_cast<const ulonglong2*>(src + m * a.n_chan);
            xa[r] = v.x; xb[r] = v.y;
        }
    }
    // ---- detrending statistics over the window (scipy.signal.detrend constant / linear; zeros included) ----
    const float tmid = 0.5f * (float)(n_win - 1);
    c2 ma = 0ull, mb = 0ull, sla = 0ull, slb = 0ull;
    const bool sloped = a.polyremoval == 1 && n_win > 1;
    auto tile_sum = [&](float (&v)[8], int nv) {                     // sum over the tile's threads with equal h
#pragma unroll
        for (int off = 2; off < 32; off <<= 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i < nv) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
        }
        __syncthreads();                                             // previous readers of `red` are done
        if (lane < 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) red[(wt * 2 + lane) * 8 + i] = v[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) s += red[(w * 2 + h) * 8 + i];
            v[i] = s;
        }
    };
    if (a.polyremoval >= 0) {
        c2 sa = 0ull, sb = 0ull, ta = 0ull, tb = 0ull;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int n = o + 64 * r;
            if (n < n_win) {
                const float t = (float)n - tmid;
                sa = add2(sa, xa[r]); sb = add2(sb, xb[r]);
                ta = fma2(bc(t), xa[r], ta); tb = fma2(bc(t), xb[r], tb);
            }
        }
        float v[8] = {re(sa), im(sa), re(sb), im(sb), re(ta), im(ta), re(tb), im(tb)};
        tile_sum(v, sloped ? 8 : 4);
        const float inv_n = 1.f / (float)n_win;
        ma = pk(v[0] * inv_n, v[1] * inv_n); mb = pk(v[2] * inv_n, v[3] * inv_n);
        if (sloped) {
            const float stt = (float)((double)n_win * ((double)n_win * n_win - 1.0) / 12.0);
            sla = pk(v[4] / stt, v[5] / stt); slb = pk(v[6] / stt, v[7] / stt);
        }
    }
    // detrended frame (zero past the window): what every taper starts from
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int n = o + 64 * r;
        if (n < n_win) {
            const float t = (float)n - tmid;
            xa[r] = sub2(xa[r], fma2(sla, bc(t), ma));
            xb[r] = sub2(xb[r], fma2(slb, bc(t), mb));
        } else {
            xa[r] = 0ull; xb[r] = 0ull;
        }
    }

    // ---- per-thread addresses ----
    const uint32_t p1 = (uint32_t)(swz(o) * ROW_BYTES + h * 16);                         // + q * 64 rows
    const int g2 = o >> 3, o2 = o & 7;
    // pass 2: i = g2*64 + 8 r + o2 -> swz = i ^ (r & 3)
    const uint32_t p2 = (uint32_t)((g2 * 64 + o2) * ROW_BYTES + h * 16);                 // ^ (r&3)<<5, + r*256
    // pass 3: i = 8 o + r -> swz = i ^ (o & 3)
    const uint32_t cu = (uint32_t)(o & 3);
    const uint32_t p3 = (uint32_t)(o * 8 * ROW_BYTES + h * 16);                          // + ((r ^ cu) & 7 ...) see below
    // epilogue: thread (s, m, h) with m = k1 + 8 k2 (k2 < 4): bins k = k0 + 8 m
    const int es = o >> 5, em = o & 31, k1 = em & 7, k2 = em >> 3;
    auto slot_addr = [&](int d0, int d1, int d2) {       // bin with digits (d0, d1, d2) sits at pos d0*64 + d1*8 + d2
        const int pos = d0 * 64 + d1 * 8 + d2;
        return (uint32_t)(swz(pos) * ROW_BYTES + h * 16);
    };

    const c2 hs = bc(0.5f * a.scale);
    const float inv_ntap = 1.f / (float)K;
    const bool accumulate = !a.keeptapers && K > 1;
    // running taper sums of this thread's 4 group items (+ Nyquist for 2 threads): KIND 0: 4 floats, else 8
    float acc[5][KIND == 0 ? 4 : 8];
#pragma unroll
    for (int g = 0; g < 5; ++g)
#pragma unroll
        for (int i = 0; i < (KIND == 0 ? 4 : 8); ++i) acc[g][i] = 0.f;

    for (int k = 0; k < K; ++k) {
        const float* __restrict__ wk = wins + k * N;
        c2 ya[8], yb[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const c2 w = bc(wk[o + 64 * r]);
            ya[r] = mul2(xa[r], w);
            yb[r] = mul2(xb[r], w);
        }
        if (a.demean_taper) {                                        // mtmfft.py:114-116
            c2 sa = 0ull, sb = 0ull;
#pragma unroll
            for (int r = 0; r < 8; ++r) { sa = add2(sa, ya[r]); sb = add2(sb, yb[r]); }
            float v[8] = {re(sa), im(sa), re(sb), im(sb), 0.f, 0.f, 0.f, 0.f};
            tile_sum(v, 4);
            const float inv_n = 1.f / (float)n_win;
            const c2 tma = pk(v[0] * inv_n, v[1] * inv_n), tmb = pk(v[2] * inv_n, v[3] * inv_n);
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (o + 64 * r < n_win) { ya[r] = sub2(ya[r], tma); yb[r] = sub2(yb[r], tmb); }
        }
        // pass 1 (stride 64): registers -> butterflies -> shared memory
        dft8(ya);
        dft8(yb);
        __syncthreads();                                             // the previous taper's epilogue reads are done
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            c2 va = ya[q], vb = yb[q];
            if (q > 0) { const float2 w = tw1[(q - 1) * 64 + o]; va = cmul2(va, w.x, w.y); vb = cmul2(vb, w.x, w.y); }
            *reinterpret_cast<ulonglong2*>(work + p1 + q * 64 * ROW_BYTES) = make_ulonglong2(va, vb);
        }
        __syncthreads();
        // pass 2 (stride 8)
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(work + ((p2 ^ (uint32_t)((r & 3) << 5)) + r * 8 * ROW_BYTES));
            ya[r] = v.x; yb[r] = v.y;
        }
        dft8(ya);
        dft8(yb);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            c2 va = ya[q], vb = yb[q];
            if (q > 0) { const float2 w = tw2[(q - 1) * 8 + o2]; va = cmul2(va, w.x, w.y); vb = cmul2(vb, w.x, w.y); }
            *reinterpret_cast<ulonglong2*>(work + ((p2 ^ (uint32_t)((q & 3) << 5)) + q * 8 * ROW_BYTES)) = make_ulonglong2(va, vb);
        }
        __syncthreads();
        // pass 3 (stride 1): slot of element r is 8 o + (r ^ (o & 3)) (the XOR touches the low two bits of r only)
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(work + p3 + (((uint32_t)r ^ cu) << 5));
            ya[r] = v.x; yb[r] = v.y;
        }
        dft8(ya);
        dft8(yb);
#pragma unroll
        for (int q = 0; q < 8; ++q)
            *reinterpret_cast<ulonglong2*>(work + p3 + (((uint32_t)q ^ cu) << 5)) = make_ulonglong2(ya[q], yb[q]);
        __syncthreads();

        // ---- epilogue: split the pairs, scale, convert, accumulate / store ----
        const long long off0 = (long long)trial * a.so_trial + (long long)frame * a.so_frame +
                               (a.keeptapers ? (long long)k * a.so_taper : 0LL) + c0 + 4 * h;
        auto emit = [&](int g, uint32_t ad1, uint32_t ad2, int kf, bool mine) {
            const ulonglong2 z1 = *reinterpret_cast<const ulonglong2*>(work + ad1);
            const ulonglong2 z2 = *reinterpret_cast<const ulonglong2*>(work + ad2);
            const c2 s0 = add2(z1.x, z2.x), d0 = sub2(z1.x, z2.x), s1 = add2(z1.y, z2.y), d1 = sub2(z1.y, z2.y);
            const c2 re0 = mul2(s0, hs), im0 = mul2(mul_mi(d0), hs);     // (re X_c, re X_c+1), (im X_c, im X_c+1)
            const c2 re1 = mul2(s1, hs), im1 = mul2(mul_mi(d1), hs);
            float v[KIND == 0 ? 4 : 8];
            if constexpr (KIND == 0) {
                v[0] = convert_real(make_float2(re(re0), re(im0)), a.out_kind);
                v[1] = convert_real(make_float2(im(re0), im(im0)), a.out_kind);
                v[2] = convert_real(make_float2(re(re1), re(im1)), a.out_kind);
                v[3] = convert_real(make_float2(im(re1), im(im1)), a.out_kind);
            } else {
                v[0] = re(re0); v[1] = re(im0); v[2] = im(re0); v[3] = im(im0);
                v[4] = re(re1); v[5] = re(im1); v[6] = im(re1); v[7] = im(im1);
            }
            if (accumulate) {
#pragma unroll
                for (int i = 0; i < (KIND == 0 ? 4 : 8); ++i) { acc[g][i] += v[i]; v[i] = acc[g][i] * inv_ntap; }
                if (k < K - 1) return;
            }
            if (!mine || !active) return;
            const long long off = off0 + (long long)kf * a.so_freq;
            if constexpr (KIND == 0) {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + off) = make_float4(v[0], v[1], v[2], v[3]);
            } else if constexpr (KIND == 1) {
                float2* op = reinterpret_cast<float2*>(a.out) + off;
                *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(op + 2) = make_float4(v[4], v[5], v[6], v[7]);
            } else {
                float* op = reinterpret_cast<float*>(a.out) + off;
                *reinterpret_cast<float4*>(op) = make_float4(v[0], v[2], v[4], v[6]);
                *reinterpret_cast<float4*>(op + a.n_chan) = make_float4(v[1], v[3], v[5], v[7]);
            }
        };
#pragma unroll
        for (int g = 1; g <= 3; ++g) {
            // bins k = k0 + 8 m, k0 = g (s = 0) or 8 - g (s = 1); partner N - k has digits (8 - k0, 7 - k1, 7 - k2)
            const int k0 = es ? 8 - g : g;
            emit(g, slot_addr(k0, k1, k2), slot_addr(8 - k0, 7 - k1, 7 - k2), k0 + 8 * em, true);
        }
        {
            // k0 in {0, 4} (self-paired blocks): partner (N - k) mod N
            const int kf = 4 * es + 8 * em;
            const int kn = (N - kf) & (N - 1);
            emit(0, slot_addr(4 * es, k1, k2), slot_addr(kn & 7, (kn >> 3) & 7, kn >> 6), kf, true);
            // Nyquist bin k = 256: digits (0, 0, 4), its own partner; threads tt < 2 own it
            emit(4, slot_addr(0, 0, 4), slot_addr(0, 0, 4), N / 2, tt < 2);
        }
    }
}

template <int KIND>
int launch(const R8Args& ra, cudaStream_t stream) {
    auto kern = mtm_r8_kernel<KIND>;
    const size_t smem = (size_t)TILES * TILE_BYTES + (size_t)ra.m.n_tapers * N * 4 + TILES * 4 * 2 * 8 * 4 + (7 * 64 + 7 * 8) * 8;
    SPYB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long blocks = (ra.n_tiles + TILES - 1) / TILES;
    if (blocks > 2147483647LL) return fail("mtm_r8: too many tiles");
    kern<<<(unsigned)blocks, THREADS, smem, stream>>>(ra);
    SPYB_LAUNCH_CHECK("mtm_r8_kernel");
    count_launch();
    return 0;
}

}  // namespace

// Returns -1 when the shape is not handled here (the caller then uses mtm_dif.cu / mtm.cu).
int mtm_launch_r8(int log2n, const MtmArgs& a, cudaStream_t stream) {
    static const bool disabled = getenv("SPYB_MTM_NO_R8") != nullptr;
    if (disabled || log2n != 9) return -1;
    if (a.n_chan % 8 != 0 || !a.vec16) return -1;
    if (a.freq_idx != nullptr || a.n_freq_out != N / 2 + 1) return -1;
    if (a.chan_amax != nullptr || a.n_tapers > MAX_TAPERS_SMEM) return -1;
    if (a.n_samples < 1 || a.n_trials < 1 || a.n_frames < 1) return -1;
    const bool cplx = a.out_kind == OUT_FOURIER;
    const long long al = cplx ? 2 : 4;
    if (a.so_trial % al || a.so_frame % al || a.so_taper % al || a.so_freq % al) return -1;
    if (reinterpret_cast<uintptr_t>(a.out) % 16 != 0) return -1;
    if (!a.keeptapers && a.out_kind == OUT_FOURIER_PLANAR) return -1;
    R8Args ra;
    ra.m = a;
    ra.chan_tiles = a.n_chan / 8;
    ra.n_tiles = (long long)a.n_trials * a.n_frames * ra.chan_tiles;
    if (a.out_kind == OUT_FOURIER) return launch<1>(ra, stream);
    if (a.out_kind == OUT_FOURIER_PLANAR) return launch<2>(ra, stream);
    return launch<0>(ra, stream);
}

}  // namespace spyb
