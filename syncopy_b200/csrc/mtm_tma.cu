// K1 / K4, pipelined variant for N = 4096: persistent blocks, TMA tile loads and TMA result stores overlapped with
// the FFT passes, packed-FP32 (FADD2 / FMUL2 / FFMA2) butterflies on two channel pairs per thread.
//
// Same arithmetic as mtm_dif.cu (detrend -> taper -> [de-mean] -> real FFT of two channels packed as one complex
// series -> scale -> convert -> store; replaces syncopy/specest/mtmfft.py:111-127 and the frame loop of
// syncopy/specest/stft.py:119-154), different execution shape:
//
//   * a work item is a tile = 8 channels (4 complex pairs, 32-byte rows) x 4096 samples of one frame
//     = 128 KB of shared memory, transformed in place by three radix-16 DIF passes;
//   * 512 threads; thread (o, h) owns butterfly o of every pass and the 16-byte half h of its 32-byte slots, i.e.
//     TWO complex pairs: every shared-memory access is a 128-bit one, and a complex add / multiply on a pair is one /
//     two packed instructions (add.f32x2, mul.f32x2 + fma.f32x2 with ptxas folding the re<->im swaps and the
//     negations into operand modifiers);
//   * the raw tile arrives by TMA (cp.async.bulk.tensor, 16 boxes of 256 rows x 32 bytes, zero fill outside the
//     trial = the zero padding of stft.py:101-117), in natural row order.  The first pass reads ALL of it into
//     registers (one butterfly per thread: the whole tile sits in the register file for a moment), which
//       - makes the detrending sums free (they are taken from the registers, block-reduced across the barrier
//         that the layout change needs anyway),
//       - lets the pass write its results in the bank-conflict-free swizzled layout of the later passes, and
//       - frees the staging blocks: the TMA loads of the NEXT work item are issued right there and overlap the
//         butterflies, passes 2 and 3 and the epilogue;
//   * 8 of the 16 row blocks of a tile land in a staging area next to the work buffer, the others in place: the
//     epilogue walks the spectrum in groups k mod 16 in {a, 16 - a}, which touch exactly the two 256-slot blocks
//     a and 16 - a of the (digit-reversed) work buffer, so the in-place blocks are released -- and re-filled by TMA
//     for the next item -- while the rest of the epilogue is still running (measured: the loads are fully hidden,
//     0.374 ms with result stores disabled against 0.338 ms with loads disabled as well);
//   * results leave as 16-byte stores of two adjacent lanes = one 32-byte sector per (bin, plane).  Measured on
//     B200 (tools/micro/scatter_store2.cu): writing the 839 MB of cfg-2 spectra as 32-byte pieces 400 KB apart
//     takes 0.305 ms with STG.128 and 0.72 ms through TMA tensor stores (one box = 128 bins x 2 planes), so the
//     staged TMA-store epilogue that was tried here (0.90 ms per launch) was dropped again;
//   * slot index = i ^ ((i >> 4) & 3) ^ ((i >> 8) & 3): with 16-byte accesses a quarter warp (8 lanes) must cover
//     the eight 16-byte columns of a 128-byte line; 2 XOR bits per radix-16 digit do that for the three pass
//     strides and for the digit-reversed epilogue reads;
//   * the first pass's twiddles W^(o q) are products of four table entries W^o, W^2o, W^4o, W^8o that a thread
//     keeps in registers for the whole launch (o is fixed per thread), the window lives in shared memory: no
//     global loads besides the second pass's 2 KB twiddle table in the steady state of a single-taper launch.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "common.cuh"
#include "mtm_args.cuh"
#include "packed.cuh"
#include "spyb_internal.h"

namespace spyb {
namespace {

// ---- mbarrier / TMA wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Bounded wait: a pipeline bug must surface as a launch error (trap) within seconds, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint64_t t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P1;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if ((spin & 1023u) == 1023u) {
            const uint64_t now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 8000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2_) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2_) : "memory");
}


// ---- geometry ------------------------------------------------------------------------------------------------------
constexpr int LOG2N = 12;
constexpr int N = 1 << LOG2N;            // 4096 samples
constexpr int THREADS = N / 8;           // 512: one radix-16 butterfly (of 2 pairs) per thread and pass
constexpr int BLK_ROWS = N / 16;         // 256 rows per block = stride of the first pass
constexpr int ROW_BYTES = 32;            // 8 channels = 4 complex pairs
constexpr int BLK_BYTES = BLK_ROWS * ROW_BYTES;      // 8 KB
constexpr int WORK_BYTES = 16 * BLK_BYTES;           // 128 KB
// Epilogue group order: a = 1, 2, ..., 7 (blocks a and 16 - a), then the self-paired blocks 0 and 8.  The blocks of
// groups 1 .. NIPG are loaded in place (released early by the epilogue), all other blocks are staged.
constexpr int NST = 8;
constexpr int NIPG = (16 - NST) / 2;                 // in-place groups
constexpr int NBATCH_IP = (NIPG + 1) / 2;            // in-place blocks are re-filled two groups at a time
__host__ __device__ constexpr bool staged(int r) { return r == 0 || r == 8 || (r < 8 ? r > NIPG : (16 - r) > NIPG); }
__host__ __device__ constexpr int stage_idx(int r) {
    int n = 0;
    for (int i = 0; i < r; ++i) n += staged(i) ? 1 : 0;
    return n;
}
__host__ __device__ constexpr int raw_off(int r) {   // byte offset of raw block r in shared memory
    return staged(r) ? WORK_BYTES + stage_idx(r) * BLK_BYTES : r * BLK_BYTES;
}
constexpr int WIN_OFF = WORK_BYTES + NST * BLK_BYTES;       // float[N]: current taper, zero past the window
constexpr int TW1_OFF = WIN_OFF + N * 4;                    // float2[4][256]: W_4096^(o q), q = 1, 2, 4, 8
constexpr int TW2_OFF = TW1_OFF + 4 * BLK_ROWS * 8;         // float2[15][16]: W_256^(o' q), q = 1..15
constexpr int RED_OFF = TW2_OFF + 15 * 16 * 8;              // [16 warps][2 halves][8 floats]
constexpr int RED2_OFF = RED_OFF + 16 * 2 * 8 * 4;          // [16 warps][2 halves][4 floats]
constexpr int BAR_OFF = RED2_OFF + 16 * 2 * 4 * 4;
constexpr int SMEM_BYTES = BAR_OFF + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct TmaArgs {
    MtmArgs m;
    int dbg;               // experiments only (SPYB_MTM_DBG): 1 skip result stores, 4 no TMA loads
    int chan_tiles;
    long long n_tiles;     // trials x frames x channel tiles
};

// PLANAR: result as two float32 planes (out_kind 8), the layout the cross-spectral kernel consumes
template <bool PLANAR>
__global__ void __launch_bounds__(THREADS, 1) mtm_tma_kernel(const __grid_constant__ CUtensorMap tmap, const TmaArgs ta) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const MtmArgs& a = ta.m;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h = tid & 1;
    const int o = tid >> 1;                       // butterfly index of every pass, 0..255
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar = sbase + BAR_OFF;
    float* red = reinterpret_cast<float*>(smem + RED_OFF);
    float* red2 = reinterpret_cast<float*>(smem + RED2_OFF);
    float* wins = reinterpret_cast<float*>(smem + WIN_OFF);

    if ((long long)blockIdx.x >= ta.n_tiles) return;
    // this block's work items = tiles blockIdx.x, blockIdx.x + gridDim.x, ... (single-taper launches only: with
    // several tapers mtm_dif.cu, which keeps re-reading the tile from L2, measured faster)
    const unsigned my_items = (unsigned)((ta.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

    struct Coords { int c0, row0, trial, frame; };
    auto tile_coords = [&](unsigned jt) {
        const unsigned tile = blockIdx.x + jt * gridDim.x;
        const unsigned ct = tile % (unsigned)ta.chan_tiles, tf = tile / (unsigned)ta.chan_tiles;
        Coords c;
        c.frame = (int)(tf % (unsigned)a.n_frames);
        c.trial = (int)(tf / (unsigned)a.n_frames);
        c.c0 = (int)ct * 8;
        c.row0 = a.frame_start0 + c.frame * a.hop;
        return c;
    };
    auto issue_staged = [&](const Coords& c) {
        if (ta.dbg & 4) return;
        mbar_expect_tx(bar, NST * BLK_BYTES);
#pragma unroll
        for (int r = 0; r < 16; ++r)
            if (staged(r)) tma_load_3d(&tmap, bar, sbase + raw_off(r), c.c0, c.row0 + r * BLK_ROWS, c.trial);
    };
    auto issue_inplace = [&](const Coords& c, int batch) {       // groups 2*batch + 1, 2*batch + 2
        if (ta.dbg & 4) return;
        int n = 0;
#pragma unroll
        for (int g = 2 * batch + 1; g <= 2 * batch + 2 && g <= NIPG; ++g) n += 2;
        mbar_expect_tx(bar, n * BLK_BYTES);
#pragma unroll
        for (int g = 2 * batch + 1; g <= 2 * batch + 2 && g <= NIPG; ++g) {
            tma_load_3d(&tmap, bar, sbase + raw_off(g), c.c0, c.row0 + g * BLK_ROWS, c.trial);
            tma_load_3d(&tmap, bar, sbase + raw_off(16 - g), c.c0, c.row0 + (16 - g) * BLK_ROWS, c.trial);
        }
    };
    auto load_window = [&](int k) {      // taper k -> shared memory, zero past the window (8 floats per thread)
        const float* __restrict__ w = a.tapers + (long long)k * a.n_win;
#pragma unroll
        for (int i = 0; i < N / THREADS; ++i) {
            const int n = tid + i * THREADS;
            wins[n] = n < a.n_win ? __ldg(w + n) : 0.f;
        }
    };

    if (tid == 0) {
        mbar_init(bar, 1 + NBATCH_IP);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    Coords cur = tile_coords(0);
    if (tid == 0) {
        issue_staged(cur);
#pragma unroll
        for (int b = 0; b < NBATCH_IP; ++b) issue_inplace(cur, b);
    }
    load_window(0);

    const int n_win = a.n_win;
    const float tmid = 0.5f * (float)(n_win - 1);
    const bool full_win = n_win == N;
    const float half_scale = 0.5f * a.scale;
    // twiddle tables -> shared memory (the steady state then has no global or local loads at all: a load queued
    // behind the result stores of the previous item would expose their whole drain time)
    float2* stw1 = reinterpret_cast<float2*>(smem + TW1_OFF);
    float2* stw2 = reinterpret_cast<float2*>(smem + TW2_OFF);
    for (int i = tid; i < 4 * BLK_ROWS; i += THREADS) {
        const int qi = i >> 8, oo = i & 255;                 // q = 1 << qi, table entry q at [(q-1)*256 + o]
        stw1[i] = __ldg(a.tw_dif + ((1 << qi) - 1) * BLK_ROWS + oo);
    }
    for (int i = tid; i < 15 * 16; i += THREADS) stw2[i] = __ldg(a.tw_dif + 15 * BLK_ROWS + i);

    // per-thread shared-memory addresses (bytes from the start of the work buffer)
    const uint32_t raw_lane = (uint32_t)(o * ROW_BYTES + h * 16);
    const uint32_t p1_a0 = (uint32_t)(((o ^ ((o >> 4) & 3)) * ROW_BYTES) + h * 16);               // + q*8 KB, ^ (q&3)<<5
    const int g2 = o >> 4, o2 = o & 15;
    const uint32_t p2_a0 = (uint32_t)(((g2 * 256 + (o2 ^ (g2 & 3))) * ROW_BYTES) + h * 16);       // + r*512, ^ (r&3)<<5
    const uint32_t cu = (uint32_t)((o & 3) ^ ((o >> 4) & 3));
    const uint32_t p3_a0 = (uint32_t)(o * 16 * ROW_BYTES + h * 16);                               // + (r>>2)*128 + ((r&3)^cu)<<5
    // epilogue: thread (m, s, h), bin k = k0 + 16 m with k1 = m & 15, k2 = m >> 4 (< 8)
    const int em = o & 127, es = o >> 7;
    const int k1 = em & 15, k2 = em >> 4;
    const uint32_t e_t1 = (uint32_t)(((k1 * 16 + (k2 ^ (k1 & 3))) * ROW_BYTES) + h * 16);
    const uint32_t e_t2 = (uint32_t)((((15 - k1) * 16 + ((15 - k2) ^ ((15 - k1) & 3))) * ROW_BYTES) + h * 16);

    for (unsigned j = 0; j < my_items; ++j) {
        const bool has_next = j + 1 < my_items;
        Coords nxt = cur;
        if (has_next) nxt = tile_coords(j + 1);
        const int c0 = cur.c0, trial = cur.trial, frame = cur.frame;

        if (!(ta.dbg & 4)) mbar_wait(bar, (uint32_t)(j & 1));

        // ---------------- pass 1: raw rows -> registers ----------------
        c2 xa[16], xb[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(smem + raw_off(r) + raw_lane);
            xa[r] = v.x; xb[r] = v.y;
        }
        // detrending sums over the window (scipy.signal.detrend: constant / linear), zeros outside the trial included
        c2 ma = 0ull, mb = 0ull, sla = 0ull, slb = 0ull;     // mean / slope of the thread's 4 channels (packed pairs)
        if (a.polyremoval >= 0) {
            c2 sa = 0ull, sb = 0ull, ta_ = 0ull, tb_ = 0ull;
            if (full_win && a.polyremoval == 0) {
#pragma unroll
                for (int r = 0; r < 16; ++r) { sa = add2(sa, xa[r]); sb = add2(sb, xb[r]); }
            } else {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const int n = o + r * BLK_ROWS;
                    if (n < n_win) {
                        const float t = (float)n - tmid;
                        sa = add2(sa, xa[r]); sb = add2(sb, xb[r]);
                        ta_ = fma2(bc(t), xa[r], ta_); tb_ = fma2(bc(t), xb[r], tb_);
                    }
                }
            }
            float v[8] = {re(sa), im(sa), re(sb), im(sb), re(ta_), im(ta_), re(tb_), im(tb_)};
            const int nv = a.polyremoval == 1 ? 8 : 4;
#pragma unroll
            for (int off = 2; off < 32; off <<= 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (i < nv) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
            }
            if (lane < 2) {
                float4* dst = reinterpret_cast<float4*>(red + (warp * 2 + lane) * 8);
                dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                if (nv == 8) dst[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
        __syncthreads();                       // every raw row is in registers; `red` is complete
        if (tid == 0 && has_next) {            // the staging area is free: fetch the next item's staged blocks
            fence_proxy_async();
            issue_staged(nxt);
        }
        if (a.polyremoval >= 0) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f), t4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int w = 0; w < 16; ++w) {     // fixed order: deterministic
                const float4 p = *reinterpret_cast<const float4*>(red + (w * 2 + h) * 8);
                s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
            }
            const float inv_n = 1.f / (float)n_win;
            ma = pk(s.x * inv_n, s.y * inv_n); mb = pk(s.z * inv_n, s.w * inv_n);
            if (a.polyremoval == 1 && n_win > 1) {
#pragma unroll
                for (int w = 0; w < 16; ++w) {
                    const float4 p = *reinterpret_cast<const float4*>(red + (w * 2 + h) * 8 + 4);
                    t4.x += p.x; t4.y += p.y; t4.z += p.z; t4.w += p.w;
                }
                const float stt = (float)((double)n_win * ((double)n_win * n_win - 1.0) / 12.0);
                sla = pk(t4.x / stt, t4.y / stt); slb = pk(t4.z / stt, t4.w / stt);
            }
        }
        const bool sloped = a.polyremoval == 1 && n_win > 1;

        // taper (x - trend) * w; rows past the window are zero
        if (full_win && !sloped) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const c2 w = bc(wins[o + r * BLK_ROWS]);
                xa[r] = mul2(sub2(xa[r], ma), w);
                xb[r] = mul2(sub2(xb[r], mb), w);
            }
        } else {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int n = o + r * BLK_ROWS;
                if (n < n_win) {
                    const float t = (float)n - tmid;
                    const c2 w = bc(wins[n]);
                    xa[r] = mul2(sub2(xa[r], fma2(sla, bc(t), ma)), w);
                    xb[r] = mul2(sub2(xb[r], fma2(slb, bc(t), mb)), w);
                } else {
                    xa[r] = 0ull; xb[r] = 0ull;
                }
            }
        }
        if (a.demean_taper) {
            // mean of the tapered window (mtmfft.py:114-116), subtracted inside the window
            c2 sa = 0ull, sb = 0ull;
#pragma unroll
            for (int r = 0; r < 16; ++r) { sa = add2(sa, xa[r]); sb = add2(sb, xb[r]); }     // rows past the window are 0
            float v[4] = {re(sa), im(sa), re(sb), im(sb)};
#pragma unroll
            for (int off = 2; off < 32; off <<= 1) {
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
            }
            if (lane < 2) *reinterpret_cast<float4*>(red2 + (warp * 2 + lane) * 4) = make_float4(v[0], v[1], v[2], v[3]);
            __syncthreads();
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int w = 0; w < 16; ++w) {
                const float4 p = *reinterpret_cast<const float4*>(red2 + (w * 2 + h) * 4);
                s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
            }
            const float inv_n = 1.f / (float)n_win;
            const c2 tma_ = pk(s.x * inv_n, s.y * inv_n), tmb_ = pk(s.z * inv_n, s.w * inv_n);
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                if (full_win || o + r * BLK_ROWS < n_win) { xa[r] = sub2(xa[r], tma_); xb[r] = sub2(xb[r], tmb_); }
            }
        }

        // butterflies + twiddles, results into the swizzled in-place layout
        dft16(xa);
        dft16(xb);
        {
            // W^(o q) for q = 1..15 from the four table entries the thread keeps (at most 3 products deep)
            c2 wq[16];
            {
                const float2 t1 = stw1[o], t2 = stw1[BLK_ROWS + o], t4 = stw1[2 * BLK_ROWS + o], t8 = stw1[3 * BLK_ROWS + o];
                wq[1] = pk(t1.x, t1.y); wq[2] = pk(t2.x, t2.y); wq[4] = pk(t4.x, t4.y); wq[8] = pk(t8.x, t8.y);
            }
            const c2 wq1 = wq[1], wq2 = wq[2], wq4 = wq[4], wq8 = wq[8];
            wq[3] = cmulc2(wq2, wq1); wq[5] = cmulc2(wq4, wq1); wq[6] = cmulc2(wq4, wq2); wq[7] = cmulc2(wq4, wq[3]);
            wq[9] = cmulc2(wq8, wq1); wq[10] = cmulc2(wq8, wq2); wq[11] = cmulc2(wq8, wq[3]); wq[12] = cmulc2(wq8, wq4);
            wq[13] = cmulc2(wq8, wq[5]); wq[14] = cmulc2(wq8, wq[6]); wq[15] = cmulc2(wq8, wq[7]);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                c2 va = xa[reg16(q)], vb = xb[reg16(q)];
                if (q > 0) { va = cmulc2(va, wq[q]); vb = cmulc2(vb, wq[q]); }
                *reinterpret_cast<ulonglong2*>(smem + ((p1_a0 ^ (uint32_t)((q & 3) << 5)) + q * BLK_BYTES)) = make_ulonglong2(va, vb);
            }
        }
        __syncthreads();

        // ---------------- pass 2: stride 16 ----------------
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(smem + ((p2_a0 ^ (uint32_t)((r & 3) << 5)) + r * 16 * ROW_BYTES));
            xa[r] = v.x; xb[r] = v.y;
        }
        dft16(xa);
        dft16(xb);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            c2 va = xa[reg16(q)], vb = xb[reg16(q)];
            if (q > 0) {
                const float2 w = stw2[(q - 1) * 16 + (o & 15)];
                va = cmul2(va, w.x, w.y);
                vb = cmul2(vb, w.x, w.y);
            }
            *reinterpret_cast<ulonglong2*>(smem + ((p2_a0 ^ (uint32_t)((q & 3) << 5)) + q * 16 * ROW_BYTES)) = make_ulonglong2(va, vb);
        }
        __syncthreads();

        // ---------------- pass 3: stride 1 ----------------
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(
                smem + (p3_a0 + ((((uint32_t)(r & 3)) ^ cu) << 5) + (r >> 2) * 128));
            xa[r] = v.x; xb[r] = v.y;
        }
        dft16(xa);
        dft16(xb);
#pragma unroll
        for (int q = 0; q < 16; ++q)
            *reinterpret_cast<ulonglong2*>(smem + (p3_a0 + ((((uint32_t)(q & 3)) ^ cu) << 5) + (q >> 2) * 128)) =
                make_ulonglong2(xa[reg16(q)], xb[reg16(q)]);
        __syncthreads();

        // ---------------- epilogue: split the pairs, scale, convert, store ----------------
        const c2 hs = bc(half_scale);
        const long long off0 = (long long)trial * a.so_trial + (long long)frame * a.so_frame +
                               c0 + 4 * h;
        auto emit = [&](uint32_t ad1, uint32_t ad2, int kf) {
            const ulonglong2 z1 = *reinterpret_cast<const ulonglong2*>(smem + ad1);
            const ulonglong2 z2 = *reinterpret_cast<const ulonglong2*>(smem + ad2);
            // pair (c, c+1): X_c = (z1 + conj z2)/2, X_{c+1} = (z1 - conj z2)/(2i)
            const c2 s0 = add2(z1.x, z2.x), d0 = sub2(z1.x, z2.x), s1 = add2(z1.y, z2.y), d1 = sub2(z1.y, z2.y);
            const c2 re0 = mul2(s0, hs), im0 = mul2(mul_mi(d0), hs);     // (re X_c, re X_{c+1}), (im X_c, im X_{c+1})
            const c2 re1 = mul2(s1, hs), im1 = mul2(mul_mi(d1), hs);
            if ((ta.dbg & 1) && re(re0) != 12345.678f) return;
            const long long off = off0 + (long long)kf * a.so_freq;
            if constexpr (PLANAR) {
                float* op = reinterpret_cast<float*>(a.out) + off;
                if (ta.dbg & 16) {          // experiment: streaming (evict-first) stores
                    __stcs(reinterpret_cast<float4*>(op), make_float4(re(re0), im(re0), re(re1), im(re1)));
                    __stcs(reinterpret_cast<float4*>(op + a.n_chan), make_float4(re(im0), im(im0), re(im1), im(im1)));
                    return;
                }
                *reinterpret_cast<ulonglong2*>(op) = make_ulonglong2(re0, re1);
                *reinterpret_cast<ulonglong2*>(op + a.n_chan) = make_ulonglong2(im0, im1);
            } else if (a.out_kind == OUT_FOURIER) {
                float2* op = reinterpret_cast<float2*>(a.out) + off;
                *reinterpret_cast<float4*>(op) = make_float4(re(re0), re(im0), im(re0), im(im0));
                *reinterpret_cast<float4*>(op + 2) = make_float4(re(re1), re(im1), im(re1), im(im1));
            } else {
                float* op = reinterpret_cast<float*>(a.out) + off;
                float4 r4;
                r4.x = convert_real(make_float2(re(re0), re(im0)), a.out_kind);
                r4.y = convert_real(make_float2(im(re0), im(im0)), a.out_kind);
                r4.z = convert_real(make_float2(re(re1), re(im1)), a.out_kind);
                r4.w = convert_real(make_float2(im(re1), im(im1)), a.out_kind);
                *reinterpret_cast<float4*>(op) = r4;
            }
        };
#pragma unroll
        for (int g = 1; g <= 7; ++g) {
            // bins k = k0 + 16 m, k0 = g (s = 0) or 16 - g (s = 1); partner N - k sits in block 16 - k0
            const int k0 = es ? 16 - g : g;
            const uint32_t x1 = (uint32_t)((k0 & 3) << 5), x2 = (uint32_t)(((16 - k0) & 3) << 5);
            emit((uint32_t)(k0 * BLK_BYTES) + (e_t1 ^ x1), (uint32_t)((16 - k0) * BLK_BYTES) + (e_t2 ^ x2), k0 + 16 * em);
            if (g <= NIPG && (g % 2 == 0 || g == NIPG)) {
                // blocks of groups (g-1, g) are consumed: refill them for the next item while the epilogue goes on
                __syncthreads();
                if (tid == 0 && has_next) {
                    fence_proxy_async();
                    issue_inplace(nxt, (g - 1) / 2);
                }
            }
        }
        {
            // blocks 0 and 8 (self-paired): k = 8 s + 16 m, partner (N - k) mod N in the same block
            const int kf = 8 * es + 16 * em;
            const int kn = (N - kf) & (N - 1);
            const int q0 = kn & 15, q1 = (kn >> 4) & 15, q2 = kn >> 8;
            const uint32_t ad1 = (uint32_t)(8 * es * BLK_BYTES) + e_t1;                 // (k0 & 3) == 0
            const uint32_t ad2 = (uint32_t)(((q0 * 256 + q1 * 16 + (q2 ^ (q1 & 3) ^ (q0 & 3))) * ROW_BYTES) + h * 16);
            emit(ad1, ad2, kf);
            if (tid < 2) {
                // Nyquist bin k = N/2: digits (0, 0, 8), slot 8, its own partner
                const uint32_t adn = (uint32_t)(8 * ROW_BYTES + h * 16);
                emit(adn, adn, N / 2);
            }
        }
        // no barrier here: the next item's first pass only READS shared memory before its own barrier
        cur = nxt;
    }
}

// ---- host side -------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

template <bool PLANAR>
int launch(const CUtensorMap& tmap, const TmaArgs& ta, int grid, cudaStream_t stream) {
    auto kern = mtm_tma_kernel<PLANAR>;
    // per device / context attribute: set on every launch (several engines may live in one process)
    SPYB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    kern<<<grid, THREADS, SMEM_BYTES, stream>>>(tmap, ta);
    SPYB_LAUNCH_CHECK("mtm_tma_kernel");
    count_launch();
    return 0;
}

}  // namespace

// Returns -1 when the shape is not handled here (the caller then uses mtm_dif.cu / mtm.cu).
int mtm_launch_tma(int log2n, const MtmArgs& a, cudaStream_t stream) {
    static const bool disabled = getenv("SPYB_MTM_NO_TMA") != nullptr;
    if (disabled || log2n != LOG2N) return -1;
    if (a.n_chan % 8 != 0 || !a.vec16) return -1;
    if (a.freq_idx != nullptr || a.n_freq_out != N / 2 + 1) return -1;
    if (a.n_tapers != 1) return -1;
    if (a.chan_amax != nullptr) return -1;
    if (a.n_samples < 1 || a.n_trials < 1) return -1;
    if ((long long)a.n_trials * a.n_frames * (a.n_chan / 8) >= (1LL << 31)) return -1;
    // 16-byte stores: element offsets of (frequency, trial, frame, taper) must keep 4 real / 2 complex elements aligned
    const bool cplx = a.out_kind == OUT_FOURIER;
    const long long al = cplx ? 2 : 4;
    if (a.so_trial % al || a.so_frame % al || a.so_taper % al || a.so_freq % al) return -1;
    if (reinterpret_cast<uintptr_t>(a.out) % 16 != 0) return -1;
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) return -1;

    CUtensorMap tmap;
    static const int promo = getenv("SPYB_MTM_L2PROMO") ? atoi(getenv("SPYB_MTM_L2PROMO")) : 2;
    const CUtensorMapL2promotion l2promo = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE :
                                           promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B :
                                           promo == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    const cuuint64_t gdim[3] = {(cuuint64_t)a.n_chan, (cuuint64_t)a.n_samples, (cuuint64_t)a.n_trials};
    const cuuint64_t gstride[2] = {(cuuint64_t)a.n_chan * 4, (cuuint64_t)a.trial_stride * 4};
    const cuuint32_t box[3] = {8, (cuuint32_t)BLK_ROWS, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(a.x), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (mtm input) failed with code %d", (int)r);

    TmaArgs ta;
    ta.m = a;
    static const int dbg = getenv("SPYB_MTM_DBG") ? atoi(getenv("SPYB_MTM_DBG")) : 0;
    ta.dbg = dbg;
    ta.chan_tiles = a.n_chan / 8;
    ta.n_tiles = (long long)a.n_trials * a.n_frames * ta.chan_tiles;
    int dev = 0, n_sm = 148;
    SPYB_CUDA(cudaGetDevice(&dev));
    SPYB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    const int grid = (int)(ta.n_tiles < n_sm ? ta.n_tiles : n_sm);
    return a.out_kind == OUT_FOURIER_PLANAR ? launch<true>(tmap, ta, grid, stream) : launch<false>(tmap, ta, grid, stream);
}

}  // namespace spyb
