// K2 (tensor-core variant): cross-spectral contraction on tcgen05 with TMEM accumulators.
//
// Replaces the broadcasting outer product + taper mean + trial sum of
//   syncopy/connectivity/csd.py:98-102 and syncopy/shared/computational_routine.py:1022-1032
// for many (trial, taper) rows at once:
//   acc[f][i][j] = beta*acc[f][i][j] + alpha * sum_r X[f][r][i] * conj(X[f][r][j]).
//
// Input layout ("planar", written by the mtmfft kernel with out_kind = OUT_FOURIER_PLANAR):
//   float32 [f][r][plane = re|im][c], i.e. per (f, r) one row of C real parts then C imaginary parts.
// With A_p / B_p the re / im planes of one frequency ([row][channel], channel contiguous = "MN-major"
// tcgen05 operands, rows = the MMA K dimension):
//   C_re = A_re^T B_re + A_im^T B_im          C_im = A_im^T B_re - A_re^T B_im
// Every real product runs as 3xTF32 (x = hi + lo, hi = rna_tf32(x), lo = rna_tf32(x - hi);
// hi*hi + hi*lo + lo*hi, FP32 accumulation in TMEM), which keeps the result within ~5e-7 of an
// FP32 product -- plain TF32 (10-bit mantissa) would miss the 1e-5 parity bar.
//
// One persistent CTA per SM (16 warps); a work item is (frequency, upper-triangular 128x128 output tile):
//   warp 0       TMA producer: 5-D tiled bulk-tensor loads bring KC rows x 128 channels of the re and im
//                planes of the tile's row block (and column block, if different) into shared memory,
//                directly in the swizzled canonical MN-major layout (128B span, 32B atoms)
//   warps 4..7   converter: split the landed FP32 tile into hi (in place) and lo planes
//   warp 1       MMA issuer: 12 tcgen05.mma (M=128, N=128, K=8) per 8 rows into TMEM buffer b; a commit
//                frees the stage, another one hands the buffer to the epilogue at the end of a chain
//   warps 8..15  epilogue (208 registers each via setmaxnreg): pull every finished accumulation chain out
//                of TMEM (tcgen05.ld) and sum the chains in registers; store the tile and its mirror
// Accumulators: two TMEM buffers of 256 columns (C_re | C_im), lane = output row of the tile.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <type_traits>

#include <cuda_bf16.h>

#include "common.cuh"
#include "spyb_internal.h"

namespace spyb {
namespace {

constexpr int TC_KC = 16;              // rows per pipeline stage (two K=8 MMA steps)
constexpr int TC_THREADS = 512;        // 16 warps = 4 warpgroups
constexpr int TC_CONV_THREADS = 128;   // warps 4..7
constexpr int TC_EPI_THREADS = 256;    // warps 8..15
constexpr int TC_PLANE = TC_KC * 512;  // bytes: one plane (128 channels x KC rows) of a half-tile
constexpr int TC_SLOT = 4 * TC_PLANE;  // half-tile: re_hi | im_hi | re_lo | im_lo
constexpr int TC_STAGE = 2 * TC_SLOT;  // row-block half-tile (A) + column-block half-tile (B)
constexpr int TC_STAGES = 3;
constexpr int TC_MAX_RANKS = 16;

struct TcArgs {
    int n_rows, n_freq, n_chan;
    int n_tiles;               // upper-triangular 128x128 tiles per frequency: nb (nb + 1) / 2
    int n_blk;                 // nb = ceil(n_chan / 128) (1..4); the last block may be zero-padded
    int dbg;                   // SPYB_TC_DBG: 1 = skip the result stores (timing experiments only)
    int chain_ksteps;          // pipeline stages (of TC_KC rows) accumulated in TMEM before the FP32 flush
    int store_mode;            // 0: per-thread row stores (debug), 1: shared-memory transposed, coalesced
    int rewrite_hi;            // 1: store rna_tf32(x) back as the hi operand; 0: let the MMA truncate x itself
    int bf16_cross;            // 1: the cross terms hi*lo + lo*hi run as BF16 MMAs (K = 16) on bf16 copies of hi / lo
    float alpha, beta;
    float2* acc;               // [n_freq][C][C]   (store_mode 0 / 1)
    // store_mode 2 ("tile slots"): every computed 128x128 tile goes, unmirrored, to the rank that owns its
    // frequency -- plain stores through peer-mapped pointers, so the NVLink transfer of a finished tile overlaps
    // the MMAs of the next one.  Slot layout at owner o: [src_rank][f - f_begin[o]][tile][128][128] complex64.
    float2* owner_base[TC_MAX_RANKS];
    int f_begin[TC_MAX_RANKS + 1];
    int n_owners, src_rank;
    int f_rot;                 // first frequency this rank works on (0 outside tile-slot mode)
    int n_freq_work;           // tile-slot mode: frequencies processed from f_rot on (wrapping); n_freq = all,
                               // fewer = the slab of this rank itself is left to the fused launch below
    // store_mode 3 on several ranks: the launch covers the frequencies this rank owns and the epilogue adds the
    // tiles its peers stored into the local slot buffer [n_src][n_freq][n_tiles][128][128] before normalising
    const float2* add_base;
    int n_add_src, add_skip;
    // store_mode 3 ("fused coherence", one rank, all rows in one launch): the epilogue normalises with the
    // diagonal and writes the converted coherency [n_freq][C][C] (float32, or complex64 for out_kind 2) including
    // the mirrored half; the cross-spectral matrix itself never reaches memory (csd.py:118-172 fused into the sum)
    void* coh_out;
    int out_kind;
};

// ---- PTX wrappers ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// Bounded wait: a pipeline bug must surface as a launch error (trap) within seconds, never as a hung GPU.
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint64_t t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P1;\n\t"
            "}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
        if ((spin & 1023u) == 1023u) {
            const uint64_t now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 8000000000ull) __trap();       // 8 s without progress
        }
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 rx;\n\t"
        ".reg .pred px;\n\t"
        "elect.sync rx|px, 0xFFFFFFFF;\n\t"
        "@px mov.s32 %0, 1;\n\t"
        "}" : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, void* dst,
                                            int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)),
          "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

template <int R> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }

// Shared-memory matrix descriptor, MN-major 32-bit operands.  The only swizzled layout tcgen05 accepts
// for MN-major TF32 is SWIZZLE_128B_BASE32B (32-byte chunks XOR-ed with the row index, 4-row period):
// canonical layout ((4,8,m),(4,k)) : ((1,4,LBO),(32,SBO)) in floats = 32 channels x 4 rows per 512 B
// atom; LBO = byte stride between 32-channel blocks, SBO = between 4-row groups.  It is what TMA
// produces with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B for a box whose inner dimension is 32 floats.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;     // descriptor version (Blackwell)
    d |= (uint64_t)1 << 61;     // SWIZZLE_128B_BASE32B
    return d;
}
// MN-major 16-bit operands, SWIZZLE_128B: canonical layout ((8,m),(8,k)) : ((16 B, LBO),(128 B, SBO)) -- 64 channels
// (128 B) x 8 rows per 1 KB atom, 16-byte chunks XOR-ed with the row index; LBO = stride between 64-channel
// blocks, SBO = stride between 8-row groups.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;     // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;     // SWIZZLE_128B
    return d;
}
// Instruction descriptor: BF16 x BF16 -> F32 (kind::f16), both operands MN-major, K = 16.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool neg_a) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((neg_a ? 1u : 0u) << 13) | (1u << 15) | (1u << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Instruction descriptor: TF32 x TF32 -> F32, both operands MN-major.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool neg_a) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((neg_a ? 1u : 0u) << 13) | (1u << 15) | (1u << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Work item = (frequency f, upper-triangular 128x128 output tile (ti, tj)).  A CTA enumerates its items with
// k = 0 .. item_count-1:
//   store_mode 0-2: item = blockIdx.x + k * gridDim.x over (f, t) with t fastest, so the items of one frequency are
//     neighbours in the persistent round-robin and share their operand tiles through L2; in tile-slot mode `f_rot`
//     rotates the frequency order per rank (rank r starts with the slab of owner r+1 and ends with its own), so at
//     any moment the ranks of a node write to different owners instead of all to the same one;
//   store_mode 3 (fused coherence): a CTA owns whole frequencies f = blockIdx.x + m * gridDim.x and runs their
//     tiles back to back in the order (0,0), (1,1), (0,1): both diagonals are known when the off-diagonal tile ends.
// t encodes the tile: 0 -> (0,0), 1 -> (0,1), 2 -> (1,1).
struct ItemIter {
    int n_tiles, n_blk, n_freq, f_rot, fused, count;
    __device__ __forceinline__ ItemIter(const TcArgs& a)
        : n_tiles(a.n_tiles), n_blk(a.n_blk), n_freq(a.n_freq), f_rot(a.f_rot), fused(a.store_mode == 3) {
        if (fused) {
            const int mf = ((int)blockIdx.x < n_freq) ? (n_freq - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
            count = mf * n_tiles;
        } else {
            const int n_items = a.n_freq_work * n_tiles;
            count = ((int)blockIdx.x < n_items) ? (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
        }
    }
    __device__ __forceinline__ void decode(int k, int& f, int& ti, int& tj, int& t) const {
        if (fused) {
            const int m = k / n_tiles, idx = k - m * n_tiles;
            f = (int)blockIdx.x + m * (int)gridDim.x;
            // all diagonal tiles first (their rsqrt feeds every off-diagonal tile), then the others in row-major order
            if (idx < n_blk) {
                t = tri_diag(n_blk, idx);
            } else {
                int rest = idx - n_blk;
                int bi = 0;
                while (rest >= n_blk - 1 - bi) { rest -= n_blk - 1 - bi; ++bi; }
                t = tri_diag(n_blk, bi) + 1 + rest;
            }
        } else {
            const int item = (int)blockIdx.x + k * (int)gridDim.x;
            const int fl = item / n_tiles;
            t = item - fl * n_tiles;
            f = fl + f_rot;
            if (f >= n_freq) f -= n_freq;
        }
        tri_decode(n_blk, t, ti, tj);
    }
};

__global__ void __launch_bounds__(TC_THREADS, 1)
csd_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the shared array (an integer round trip would demote every
    // access below to generic LD / ST)
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)TC_STAGES * TC_STAGE);
    uint64_t* full_raw = bars;                      // [stage]  TMA landed
    uint64_t* full_conv = bars + TC_STAGES;         // [stage]  hi / lo split done
    uint64_t* empty = bars + 2 * TC_STAGES;         // [stage]  MMAs reading the stage retired
    uint64_t* acc_full = bars + 3 * TC_STAGES;      // [2]      chain accumulated in TMEM buffer b
    uint64_t* acc_empty = acc_full + 2;             // [2]      epilogue drained TMEM buffer b
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float2* staging = reinterpret_cast<float2*>(bars + 16);       // [8 epilogue warps][32 x 16] transpose buffers, 16-byte aligned

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = a.n_chan;
    const ItemIter items(a);
    const int n_ksteps = (a.n_rows + TC_KC - 1) / TC_KC;
    const int chain_ksteps = a.chain_ksteps;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&full_raw[s], 1);
            mbar_init(&full_conv[s], TC_CONV_THREADS);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], TC_EPI_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512u);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        reg_dec<40>();
        if (warp == 0) {
            // ===== TMA producer =====
            if (elect_one()) {
                int s = 0;
                uint32_t phase = 0;
                for (int item = 0; item < items.count; ++item) {
                    int f, ti, tj, t;
                    items.decode(item, f, ti, tj, t);
                    const int n_slots = ti == tj ? 1 : 2;
                    for (int ks = 0; ks < n_ksteps; ++ks) {
                        mbar_wait(&empty[s], phase ^ 1u);
                        mbar_arrive_expect_tx(&full_raw[s], (uint32_t)(n_slots * 2 * TC_PLANE));
                        uint8_t* dst = base + (size_t)s * TC_STAGE;
                        tma_load_5d(&tmap, &full_raw[s], dst, 0, ks * TC_KC, 4 * ti, 0, f);
                        tma_load_5d(&tmap, &full_raw[s], dst + TC_PLANE, 0, ks * TC_KC, 4 * ti, 1, f);
                        if (n_slots == 2) {
                            tma_load_5d(&tmap, &full_raw[s], dst + TC_SLOT, 0, ks * TC_KC, 4 * tj, 0, f);
                            tma_load_5d(&tmap, &full_raw[s], dst + TC_SLOT + TC_PLANE, 0, ks * TC_KC, 4 * tj, 1, f);
                        }
                        if (++s == TC_STAGES) { s = 0; phase ^= 1u; }
                    }
                }
            }
        } else if (warp == 1) {
            // ===== MMA issuer =====
            constexpr uint32_t lbo = TC_KC * 128u, sbo = 512u;
            constexpr uint32_t idesc_pos = make_idesc(128, 128, false);
            constexpr uint32_t idesc_neg = make_idesc(128, 128, true);
            int s = 0;
            uint32_t phase = 0, chain = 0;               // chain counts accumulation chains over all items
            for (int item = 0; item < items.count; ++item) {
                int f, ti, tj, t;
                items.decode(item, f, ti, tj, t);
                const uint32_t b_slot = ti == tj ? 0u : (uint32_t)TC_SLOT;
                for (int ks = 0; ks < n_ksteps; ++ks) {
                    const int kc = ks % chain_ksteps;                  // position inside the accumulation chain
                    const uint32_t buf = chain & 1u;
                    if (kc == 0) {                                     // epilogue must have drained this TMEM buffer
                        mbar_wait(&acc_empty[buf], ((chain >> 1) & 1u) ^ 1u);
                        tc_fence_after();
                    }
                    mbar_wait(&full_conv[s], phase);
                    tc_fence_after();
                    const bool chain_end = (kc == chain_ksteps - 1) || (ks == n_ksteps - 1);
                    if (elect_one()) {
                        const uint32_t d_re = tmem_base + buf * 256u, d_im = d_re + 128u;
                        const uint32_t st = smem_u32(base + (size_t)s * TC_STAGE);
                        if (a.bf16_cross) {
                            // slot: re_hi32 | im_hi32 | re_h16 | im_h16 | re_l16 | im_l16   (8 + 8 + 4 x 4 KB)
                            constexpr uint32_t H16 = 2 * TC_PLANE, P16 = TC_PLANE / 2;
                            constexpr uint32_t lbo16 = 2048u, sbo16 = 1024u;
                            constexpr uint32_t ib_pos = make_idesc_bf16(128, 128, false);
                            constexpr uint32_t ib_neg = make_idesc_bf16(128, 128, true);
                            const uint32_t sa = st, sb = st + b_slot;
                            const uint64_t Are_h = make_smem_desc_sw128(sa + H16, lbo16, sbo16);
                            const uint64_t Aim_h = make_smem_desc_sw128(sa + H16 + P16, lbo16, sbo16);
                            const uint64_t Are_l = make_smem_desc_sw128(sa + H16 + 2 * P16, lbo16, sbo16);
                            const uint64_t Aim_l = make_smem_desc_sw128(sa + H16 + 3 * P16, lbo16, sbo16);
                            const uint64_t Bre_h = make_smem_desc_sw128(sb + H16, lbo16, sbo16);
                            const uint64_t Bim_h = make_smem_desc_sw128(sb + H16 + P16, lbo16, sbo16);
                            const uint64_t Bre_l = make_smem_desc_sw128(sb + H16 + 2 * P16, lbo16, sbo16);
                            const uint64_t Bim_l = make_smem_desc_sw128(sb + H16 + 3 * P16, lbo16, sbo16);
                            const uint32_t first = kc > 0 ? 1u : 0u;
                            // cross terms first (small addends), K = 16 rows per instruction
                            umma_bf16(d_re, Are_l, Bre_h, ib_pos, first);
                            umma_bf16(d_re, Are_h, Bre_l, ib_pos, 1u);
                            umma_bf16(d_re, Aim_l, Bim_h, ib_pos, 1u);
                            umma_bf16(d_re, Aim_h, Bim_l, ib_pos, 1u);
                            umma_bf16(d_im, Aim_l, Bre_h, ib_pos, first);
                            umma_bf16(d_im, Aim_h, Bre_l, ib_pos, 1u);
                            umma_bf16(d_im, Are_l, Bim_h, ib_neg, 1u);
                            umma_bf16(d_im, Are_h, Bim_l, ib_neg, 1u);
#pragma unroll
                            for (int kk = 0; kk < TC_KC / 8; ++kk) {
                                const uint32_t ka = st + (uint32_t)kk * 1024u, kb = ka + b_slot;
                                const uint64_t Are_hi = make_smem_desc(ka, lbo, sbo);
                                const uint64_t Aim_hi = make_smem_desc(ka + TC_PLANE, lbo, sbo);
                                const uint64_t Bre_hi = make_smem_desc(kb, lbo, sbo);
                                const uint64_t Bim_hi = make_smem_desc(kb + TC_PLANE, lbo, sbo);
                                umma_tf32(d_re, Are_hi, Bre_hi, idesc_pos, 1u);
                                umma_tf32(d_re, Aim_hi, Bim_hi, idesc_pos, 1u);
                                umma_tf32(d_im, Aim_hi, Bre_hi, idesc_pos, 1u);
                                umma_tf32(d_im, Are_hi, Bim_hi, idesc_neg, 1u);
                            }
                        } else
#pragma unroll
                        for (int kk = 0; kk < TC_KC / 8; ++kk) {
                            const uint32_t ka = st + (uint32_t)kk * 1024u, kb = ka + b_slot;
                            const uint64_t Are_hi = make_smem_desc(ka, lbo, sbo);
                            const uint64_t Aim_hi = make_smem_desc(ka + TC_PLANE, lbo, sbo);
                            const uint64_t Are_lo = make_smem_desc(ka + 2 * TC_PLANE, lbo, sbo);
                            const uint64_t Aim_lo = make_smem_desc(ka + 3 * TC_PLANE, lbo, sbo);
                            const uint64_t Bre_hi = make_smem_desc(kb, lbo, sbo);
                            const uint64_t Bim_hi = make_smem_desc(kb + TC_PLANE, lbo, sbo);
                            const uint64_t Bre_lo = make_smem_desc(kb + 2 * TC_PLANE, lbo, sbo);
                            const uint64_t Bim_lo = make_smem_desc(kb + 3 * TC_PLANE, lbo, sbo);
                            const uint32_t first = (kc > 0 || kk > 0) ? 1u : 0u;
                            // C_re = Re^T Re + Im^T Im
                            umma_tf32(d_re, Are_lo, Bre_hi, idesc_pos, first);
                            umma_tf32(d_re, Are_hi, Bre_lo, idesc_pos, 1u);
                            umma_tf32(d_re, Aim_lo, Bim_hi, idesc_pos, 1u);
                            umma_tf32(d_re, Aim_hi, Bim_lo, idesc_pos, 1u);
                            umma_tf32(d_re, Are_hi, Bre_hi, idesc_pos, 1u);
                            umma_tf32(d_re, Aim_hi, Bim_hi, idesc_pos, 1u);
                            // C_im = Im^T Re - Re^T Im
                            umma_tf32(d_im, Aim_lo, Bre_hi, idesc_pos, first);
                            umma_tf32(d_im, Aim_hi, Bre_lo, idesc_pos, 1u);
                            umma_tf32(d_im, Are_lo, Bim_hi, idesc_neg, 1u);
                            umma_tf32(d_im, Are_hi, Bim_lo, idesc_neg, 1u);
                            umma_tf32(d_im, Aim_hi, Bre_hi, idesc_pos, 1u);
                            umma_tf32(d_im, Are_hi, Bim_hi, idesc_neg, 1u);
                        }
                        umma_commit(&empty[s]);                      // stage free once these MMAs retire
                        if (chain_end) umma_commit(&acc_full[buf]);
                    }
                    __syncwarp();
                    if (chain_end) ++chain;
                    if (++s == TC_STAGES) { s = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp < 8) {
        // ===== converter (warps 4..7): FP32 tile -> hi (in place) + lo =====
        reg_dec<56>();
        const int ct = threadIdx.x - 128;                       // 0..127
        int s = 0;
        uint32_t phase = 0;
        for (int item = 0; item < items.count; ++item) {
            int f, ti, tj, t;
            items.decode(item, f, ti, tj, t);
            const int n_slots = ti == tj ? 1 : 2;
            for (int ks = 0; ks < n_ksteps; ++ks) {
                mbar_wait(&full_raw[s], phase);
                for (int sl = 0; sl < n_slots; ++sl) {
                    float4* hi = reinterpret_cast<float4*>(base + (size_t)s * TC_STAGE + (size_t)sl * TC_SLOT);
                    float4* lo = hi + (2 * TC_PLANE) / 16;
                    if (a.bf16_cross) {
                        // fp32 planes are in the TMA layout: offset = c_blk*2048 + row*128 + c_in*4 with the 32-byte
                        // chunk index XOR-ed with row % 4; the bf16 copies go to the SWIZZLE_128B MN-major layout:
                        // blk64*2048 + (row/8)*1024 + (row%8)*128 + ((c%64/8) ^ (row%8))*16 + (c%8)*2
                        // Thread -> chunk mapping: lane bits = (16-byte chunk 0..7 | low bit of the 32-channel block |
                        // row bit 0), so a warp's bf16 stores cover two full 128-byte rows (conflict free; with 32
                        // consecutive chunks of one block they fell into half the banks, 4-way conflicts) and every
                        // address below is a per-thread constant plus a compile-time constant of the unrolled loop.
                        uint8_t* h16 = reinterpret_cast<uint8_t*>(hi) + 2 * TC_PLANE;
                        const uint32_t chunk = (uint32_t)ct & 7u, blk_lo = ((uint32_t)ct >> 3) & 1u, row_lo = ((uint32_t)ct >> 4) & 7u;
                        const uint32_t src0 = blk_lo * 2048u + row_lo * 128u + chunk * 16u;       // + (it&1)*1024 + (it&2)*2048 + (it&4)*2048
                        // un-swizzle (row % 4 is row_lo % 4 for both values of row bit 3)
                        const uint32_t lin = (chunk * 16u) ^ ((row_lo & 3u) << 5);
                        const uint32_t c_in = (lin >> 2) & 31u;
                        const uint32_t c64 = blk_lo * 32u + c_in;                                  // channel within the 64-block
                        const uint32_t dst0 = (row_lo << 7) | ((((c64 >> 3) & 7u) ^ row_lo) << 4) | ((c64 & 7u) << 1);
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const uint32_t r8 = (uint32_t)it & 1u, bh = ((uint32_t)it >> 1) & 1u, pl = (uint32_t)it >> 2;
                            float4* src = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(hi) + pl * TC_PLANE + bh * 4096u + r8 * 1024u + src0);
                            const float4 x = *src;
                            float4 h;
                            if (a.rewrite_hi) {
                                h.x = rna_tf32(x.x); h.y = rna_tf32(x.y); h.z = rna_tf32(x.z); h.w = rna_tf32(x.w);
                                *src = h;
                            } else {                             // the tensor core ignores the low 13 mantissa bits
                                h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
                                h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
                                h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
                                h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
                            }
                            const uint32_t off = bh * 2048u + r8 * 1024u + dst0;
                            __nv_bfloat162 a0 = __floats2bfloat162_rn(h.x, h.y), a1 = __floats2bfloat162_rn(h.z, h.w);
                            __nv_bfloat162 l0 = __floats2bfloat162_rn(x.x - h.x, x.y - h.y);
                            __nv_bfloat162 l1 = __floats2bfloat162_rn(x.z - h.z, x.w - h.w);
                            uint2 hv, lv;
                            hv.x = *reinterpret_cast<uint32_t*>(&a0); hv.y = *reinterpret_cast<uint32_t*>(&a1);
                            lv.x = *reinterpret_cast<uint32_t*>(&l0); lv.y = *reinterpret_cast<uint32_t*>(&l1);
                            *reinterpret_cast<uint2*>(h16 + pl * (TC_PLANE / 2) + off) = hv;
                            *reinterpret_cast<uint2*>(h16 + (2 + pl) * (TC_PLANE / 2) + off) = lv;
                        }
                        continue;
                    }
#pragma unroll 4
                    for (int q = ct; q < (2 * TC_PLANE) / 16; q += TC_CONV_THREADS) {
                        const float4 x = hi[q];
                        float4 h, l;
                        if (a.rewrite_hi) {
                            h.x = rna_tf32(x.x); h.y = rna_tf32(x.y); h.z = rna_tf32(x.z); h.w = rna_tf32(x.w);
                            hi[q] = h;
                        } else {                                 // the tensor core ignores the low 13 mantissa bits
                            h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
                            h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
                            h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
                            h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
                        }
                        l.x = rna_tf32(x.x - h.x); l.y = rna_tf32(x.y - h.y);
                        l.z = rna_tf32(x.z - h.z); l.w = rna_tf32(x.w - h.w);
                        lo[q] = l;
                    }
                }
                fence_proxy_async_smem();
                mbar_arrive(&full_conv[s]);
                if (++s == TC_STAGES) { s = 0; phase ^= 1u; }
            }
        }
    } else {
        // ===== epilogue (warps 8..15) =====
        // The tensor core adds into its FP32 accumulators with truncation, so the error of a chain grows
        // linearly with the number of MMAs feeding one accumulator (~3e-8 per MMA, measured on B200).  Chains
        // are therefore cut after `chain_ksteps` stages; the partial tile is pulled out of TMEM (double
        // buffered, so the next chain's MMAs overlap) and summed in registers with round-to-nearest FP32 adds.
        // Thread = output row i of the tile and 64 of its 128 columns: 64 re + 64 im running sums.
        reg_inc<208>();
        const int lane_grp = warp & 3;                          // TMEM lanes this warp may read
        const int chalf = (warp - 8) >> 2;                      // which 64 columns
        uint32_t chain = 0;
        for (int item = 0; item < items.count; ++item) {
            int f, ti, tj, t;
            items.decode(item, f, ti, tj, t);
            float sr[64], si[64];
#pragma unroll
            for (int c = 0; c < 64; ++c) { sr[c] = 0.f; si[c] = 0.f; }
            if (a.store_mode == 3 && a.n_add_src > 0) {
                // Several ranks: the running sums start from the partial sums of the other ranks (tile-slot layout,
                // written by their store_mode 2 launches) instead of zero -- this thread's row, its 64 columns = 512
                // contiguous bytes per source.  Issued before the first chain is waited for, so the loads fly while
                // the tensor core works on this tile.  Fixed order -> deterministic.
                const int r0 = (warp & 3) * 32, r_loc = r0 + lane, cb = ((warp - 8) >> 2) * 64;
                const bool diag_tile = ti == tj;
                // The rows sit in local HBM; a thread walks 4 chunks x (ranks - 1) sources one round trip after the
                // other, so each source's four 128-byte pieces are pulled into L2 while the source before it is being
                // added, and the first source of the NEXT tile while this tile runs (two sources per SM in flight:
                // a whole tile ahead for every SM would not fit L2 on 8 ranks).
                auto pf_rows = [&](int src, int pf_f, int pf_t, bool pf_diag) {
                    const float2* row = a.add_base + (((size_t)src * a.n_freq + pf_f) * a.n_tiles + pf_t) * (128 * 128) +
                                        (size_t)r_loc * 128 + cb;
#pragma unroll
                    for (int c0 = 0; c0 < 64; c0 += 16) {
                        if (pf_diag && cb + c0 + 15 < r0) continue;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(row + c0));
                    }
                };
                const int src_first = a.add_skip == 0 ? 1 : 0;
                if (item == 0 && src_first < a.n_add_src) pf_rows(src_first, f, t, diag_tile);
                for (int src = 0; src < a.n_add_src; ++src) {
                    if (src == a.add_skip) continue;
                    {
                        int nsrc = src + 1;
                        if (nsrc == a.add_skip) ++nsrc;
                        if (nsrc < a.n_add_src) pf_rows(nsrc, f, t, diag_tile);
                        else if (item + 1 < items.count && src_first < a.n_add_src) {
                            int nf, nti, ntj, nt;
                            items.decode(item + 1, nf, nti, ntj, nt);
                            pf_rows(src_first, nf, nt, nti == ntj);
                        }
                    }
                    // Coalesced: a warp instruction reads 4 rows x 128 bytes (8 lanes per row piece); the pieces reach
                    // the row's owner through the warp's 4 KB staging buffer (16-byte slots XOR row: conflict-free both
                    // ways).  With one row per lane a warp load touched 32 lines and the epilogue, not the tensor
                    // core, set the pace of the launch.
                    const float4* __restrict__ pwarp = reinterpret_cast<const float4*>(
                        a.add_base + (((size_t)src * a.n_freq + f) * a.n_tiles + t) * (128 * 128) + (size_t)r0 * 128 + cb);
                    float4* stg4 = reinterpret_cast<float4*>(staging + (warp - 8) * (32 * 16));
                    const int lr = lane >> 3, lq = lane & 7;
#pragma unroll
                    for (int c0 = 0; c0 < 64; c0 += 16) {
                        // pieces entirely below the diagonal are never written by the producer (nor used here)
                        if (diag_tile && cb + c0 + 15 < r0) continue;
                        float4 v[8];
#pragma unroll
                        for (int it = 0; it < 8; ++it) v[it] = __ldcs(pwarp + (size_t)(4 * it + lr) * 64 + c0 / 2 + lq);
#pragma unroll
                        for (int it = 0; it < 8; ++it) stg4[(4 * it + lr) * 8 + (lq ^ ((4 * it + lr) & 7))] = v[it];
                        __syncwarp();
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 w = stg4[lane * 8 + (q ^ (lane & 7))];
                            sr[c0 + 2 * q] += w.x; si[c0 + 2 * q] += w.y;
                            sr[c0 + 2 * q + 1] += w.z; si[c0 + 2 * q + 1] += w.w;
                        }
                        __syncwarp();
                    }
                }
            }
            for (int k0 = 0; k0 < n_ksteps; k0 += chain_ksteps, ++chain) {
                const uint32_t buf = chain & 1u;
                mbar_wait(&acc_full[buf], (chain >> 1) & 1u);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + buf * 256u + (uint32_t)(chalf * 64);
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    uint32_t vr[16], vi[16];
                    tmem_ld16(t_row + (uint32_t)c0, vr);
                    tmem_ld16(t_row + 128u + (uint32_t)c0, vi);
                    tmem_ld_wait();
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) {
                        sr[c0 + jj] += __uint_as_float(vr[jj]);
                        si[c0 + jj] += __uint_as_float(vi[jj]);
                    }
                }
                tc_fence_before();
                mbar_arrive(&acc_empty[buf]);
            }
            if (a.dbg & 1) continue;
            // ---- store: elements on / above the diagonal as they are, plus their conjugates mirrored below it;
            // nothing computed below the diagonal is used, so the result is exactly Hermitian.  A thread owns a
            // row, so the mirrored stores (fixed j, lanes = consecutive i) coalesce as they are; the direct ones
            // go through a per-warp 32 x 16 shared-memory transpose so that half-warps write 128-byte row pieces.
            const int i0 = ti * 128 + lane_grp * 32;             // first row of this warp
            const int i = i0 + lane;
            const int jb = tj * 128 + chalf * 64;
            float2* __restrict__ fmat = a.acc + (size_t)f * C * C;
            float2* stg = staging + (warp - 8) * (32 * 16);
            const bool diag_tile = ti == tj;
            if (a.store_mode == 3) {
                float* diag = reinterpret_cast<float*>(staging + 8 * 32 * 16);     // [n_blk <= 4][128] rsqrt of the diagonal
                const int r0 = lane_grp * 32, r_loc = r0 + lane, cb = chalf * 64;
                if (diag_tile) {
                    float dv = 0.f;
#pragma unroll
                    for (int c = 0; c < 64; ++c) if (cb + c == r_loc) dv = sr[c];
                    asm volatile("bar.sync 1, 256;" ::: "memory");     // nobody still reads the previous diagonal
                    if ((r_loc >> 6) == chalf) diag[ti * 128 + r_loc] = rsqrtf(dv);
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
                const float ri = diag[ti * 128 + r_loc];
                const float* __restrict__ dj = diag + tj * 128 + cb;
                const int kind = a.out_kind;
                const size_t row_dir = 0;
                // channel counts that are not multiples of 128 (but of 32): rows / columns >= C of the last block are
                // zero padding (TMA out-of-bounds fill); a warp's 32 rows and every 16- or 32-column chunk are valid
                // or padding as a whole, so the guards are warp-uniform
                if (ti * 128 + r0 >= C) continue;
                // Pointers walk with one add per element and the triangle tests are predicates on plain stores: the
                // epilogue has ~16 us per tile before it holds up the MMAs, and only two warps per scheduler.
                auto store_real = [&](auto diag_tag, auto kind_tag) {
                    constexpr bool DG = decltype(diag_tag)::value;
                    constexpr int KD = decltype(kind_tag)::value;
                    float* __restrict__ outf = reinterpret_cast<float*>(a.coh_out) + (size_t)f * C * C;
                    float* stgf = reinterpret_cast<float*>(stg);                 // [32][32] floats
#pragma unroll
                    for (int c0 = 0; c0 < 64; c0 += 32) {
                        if (DG && cb + c0 + 31 < r0) continue;                   // entirely below the diagonal
                        if (tj * 128 + cb + c0 >= C) continue;                   // padding columns
                        float* pm = outf + (size_t)(tj * 128 + cb + c0) * C + ti * 128 + r_loc;   // mirrored (j, i)
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) {
                            const int c = c0 + jj;
                            const float w = ri * dj[c];
                            const float zx = sr[c] * w;
                            float zy = si[c] * w;
                            if (DG && cb + c == r_loc) zy = 0.f;
                            float val;
                            if (KD == OUT_ABS) { const float q = zx * zx + zy * zy; val = q > 0.f ? q * rsqrtf(q) : 0.f; }
                            else if (KD == OUT_POW) val = zx * zx + zy * zy;
                            else val = convert_real(make_float2(zx, zy), kind);
                            stgf[lane * 32 + (jj ^ lane)] = val;
                            const float mv = (KD != OUT_ABS && KD != OUT_POW && (kind == OUT_IMAG || kind == OUT_ANGLE)) ? -val : val;
                            if (!DG || cb + c > r_loc) *pm = mv;
                            pm += C;
                        }
                        __syncwarp();
                        float* pd = outf + (size_t)(ti * 128 + r0) * C + tj * 128 + cb + c0 + lane;   // direct (i, j)
                        const int jcol = cb + c0 + lane;
#pragma unroll 8
                        for (int r = 0; r < 32; ++r) {
                            const float v = stgf[r * 32 + (lane ^ r)];
                            if (!DG || jcol >= r0 + r) *pd = v;
                            pd += C;
                        }
                        __syncwarp();
                    }
                };
                if (kind == OUT_FOURIER) {
                    float2* __restrict__ outc = reinterpret_cast<float2*>(a.coh_out) + (size_t)f * C * C;
#pragma unroll
                    for (int c0 = 0; c0 < 64; c0 += 16) {
                        if (diag_tile && cb + c0 + 15 < r0) continue;            // entirely below the diagonal
                        if (tj * 128 + cb + c0 >= C) continue;                   // padding columns
                        float2* pm = outc + (size_t)(tj * 128 + cb + c0) * C + ti * 128 + r_loc;
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) {
                            const int c = c0 + jj;
                            const float w = ri * dj[c];
                            float2 z = make_float2(sr[c] * w, si[c] * w);
                            if (diag_tile && cb + c == r_loc) z.y = 0.f;
                            stg[lane * 16 + (jj ^ (lane & 15))] = z;
                            if (!diag_tile || cb + c > r_loc) *pm = make_float2(z.x, -z.y);
                            pm += C;
                        }
                        __syncwarp();
#pragma unroll
                        for (int it = 0; it < 16; ++it) {
                            const int r = 2 * it + (lane >> 4), cc = lane & 15;
                            const float2 z = stg[r * 16 + (cc ^ (r & 15))];
                            if (!diag_tile || cb + c0 + cc >= r0 + r)
                                outc[(size_t)(ti * 128 + r0 + r) * C + tj * 128 + cb + c0 + cc] = z;
                        }
                        __syncwarp();
                    }
                } else if (kind == OUT_ABS) {
                    if (diag_tile) store_real(std::true_type{}, std::integral_constant<int, OUT_ABS>{});
                    else store_real(std::false_type{}, std::integral_constant<int, OUT_ABS>{});
                } else if (kind == OUT_POW) {
                    if (diag_tile) store_real(std::true_type{}, std::integral_constant<int, OUT_POW>{});
                    else store_real(std::false_type{}, std::integral_constant<int, OUT_POW>{});
                } else {
                    if (diag_tile) store_real(std::true_type{}, std::integral_constant<int, -1>{});
                    else store_real(std::false_type{}, std::integral_constant<int, -1>{});
                }
                (void)row_dir;
                continue;
            }
            if (a.store_mode == 2) {
                int o = 0;
                while (o + 1 < a.n_owners && f >= a.f_begin[o + 1]) ++o;
                const int nf_o = a.f_begin[o + 1] - a.f_begin[o];
                float2* __restrict__ tile = a.owner_base[o] +
                    (((size_t)a.src_rank * nf_o + (size_t)(f - a.f_begin[o])) * a.n_tiles + t) * (128 * 128);
                const int r0 = lane_grp * 32;                    // first tile row of this warp
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    // diagonal tiles: 16-column pieces entirely below the diagonal are never read (warp-uniform)
                    if (diag_tile && chalf * 64 + c0 + 15 < r0) continue;
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj)
                        stg[lane * 16 + (jj ^ (lane & 15))] = make_float2(sr[c0 + jj] * a.alpha, si[c0 + jj] * a.alpha);
                    __syncwarp();
#pragma unroll
                    for (int it = 0; it < 16; ++it) {
                        const int r = 2 * it + (lane >> 4), cc = lane & 15;
                        float2 o2 = stg[r * 16 + (cc ^ (r & 15))];
                        float2* dst = tile + (size_t)(r0 + r) * 128 + chalf * 64 + c0 + cc;
                        if (diag_tile && r0 + r == chalf * 64 + c0 + cc) o2.y = 0.f;    // auto-spectra are exactly real
                        if (a.beta != 0.f) { const float2 old0 = *dst; o2.x += a.beta * old0.x; o2.y += a.beta * old0.y; }
                        *dst = o2;
                    }
                    __syncwarp();
                }
                continue;
            }
            if (a.store_mode == 0) {
                float2* __restrict__ orow = fmat + (size_t)i * C;
                if (i >= C) continue;                            // padding rows
#pragma unroll
                for (int c = 0; c < 64; ++c) {
                    const int j = jb + c;
                    if (j >= C) continue;                        // padding columns
                    if (diag_tile && j < i) continue;
                    float2 o = make_float2(sr[c] * a.alpha, si[c] * a.alpha);
                    if (a.beta != 0.f) {
                        const float2 old0 = orow[j];
                        o.x += a.beta * old0.x; o.y += a.beta * old0.y;
                    }
                    if (j == i) o.y = 0.f;
                    orow[j] = o;
                    if (j != i) fmat[(size_t)j * C + i] = make_float2(o.x, -o.y);
                }
                continue;
            }
            if (i0 >= C) continue;                               // padding rows (warp-uniform: C % 32 == 0)
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
                if (diag_tile && jb + c0 + 15 < i0) continue;    // chunk entirely below the diagonal (warp-uniform)
                if (jb + c0 >= C) continue;                      // padding columns
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {
                    const int j = jb + c0 + jj;
                    float2 o = make_float2(sr[c0 + jj] * a.alpha, si[c0 + jj] * a.alpha);
                    if (a.beta != 0.f && (!diag_tile || j >= i)) {
                        const float2 old0 = fmat[(size_t)i * C + j];
                        o.x += a.beta * old0.x; o.y += a.beta * old0.y;
                    }
                    if (j == i) o.y = 0.f;                      // auto-spectra are exactly real
                    stg[lane * 16 + (jj ^ (lane & 15))] = o;
                    if (!diag_tile || j > i) fmat[(size_t)j * C + i] = make_float2(o.x, -o.y);
                }
                __syncwarp();
#pragma unroll
                for (int it = 0; it < 16; ++it) {
                    const int r = 2 * it + (lane >> 4), cc = lane & 15;
                    const int ii = i0 + r, j = jb + c0 + cc;
                    const float2 o = stg[r * 16 + (cc ^ (r & 15))];
                    if (!diag_tile || j >= ii) fmat[(size_t)ii * C + j] = o;
                }
                __syncwarp();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512u);
    }
}

// ---- host side -------------------------------------------------------------------------
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

}  // namespace

bool csd_tc_supported(int n_chan, long long sx_f, long long sx_r) {
    return n_chan >= 64 && n_chan <= 512 && n_chan % 32 == 0 && sx_r % 4 == 0 && sx_f % 4 == 0;
}

static int launch_tc(const CsdPlanarDesc& d, TcArgs& a, cudaStream_t stream) {
    if (d.n_freq <= 0 || d.n_chan <= 0) return 0;
    if (!csd_tc_supported(d.n_chan, d.sx_f, d.sx_r))
        return fail("tcgen05 CSD kernel needs 64 <= n_chan <= 512, a multiple of 32, and 16-byte aligned strides "
                    "(got n_chan=%d, sx_f=%lld, sx_r=%lld)", d.n_chan, d.sx_f, d.sx_r);
    if (reinterpret_cast<uintptr_t>(d.planes) % 16 != 0)
        return fail("tcgen05 CSD kernel needs 16-byte aligned buffers");
    if (d.n_rows <= 0) return fail("csd: n_rows must be positive");
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) return fail("cuTensorMapEncodeTiled is not available from this driver");

    const int C = d.n_chan;
    CUtensorMap tmap;
    const cuuint64_t gdim[5] = {32, (cuuint64_t)d.n_rows, (cuuint64_t)(C / 32), 2, (cuuint64_t)d.n_freq};
    const cuuint64_t gstride[4] = {(cuuint64_t)d.sx_r * 4, 128, (cuuint64_t)C * 4, (cuuint64_t)d.sx_f * 4};
    const cuuint32_t box[5] = {32, (cuuint32_t)TC_KC, 4, 1, 1};     // 128 channels x KC rows of one plane
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(d.planes), gdim, gstride, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with code %d", (int)r);

    a.n_rows = d.n_rows; a.n_freq = d.n_freq; a.n_chan = C;
    a.n_blk = (C + 127) / 128;          // boxes reaching past C are zero-filled by TMA
    a.n_tiles = tri_tiles(a.n_blk);
    a.alpha = d.alpha; a.beta = d.beta;
    // accumulation-chain length in rows (multiple of TC_KC); SPYB_TC_CHAIN_ROWS overrides for experiments
    int chain_rows = 64;
    if (const char* e = getenv("SPYB_TC_CHAIN_ROWS")) chain_rows = atoi(e);
    if (chain_rows < TC_KC) chain_rows = TC_KC;
    a.chain_ksteps = chain_rows / TC_KC;
    // hi operand: the tensor core truncates FP32 to TF32 itself, so the converter leaves the FP32 tile alone and only
    // derives lo = x - trunc(x) (measured on B200, 200 / 1400 rows: 0.715 -> 0.657 ms, error vs FP64 1.1e-6 ->
    // 2.3e-6 normwise, 3.0e-6 on the coherence scale); SPYB_TC_REWRITE_HI=1 stores round-to-nearest hi back
    a.rewrite_hi = 0;
    if (const char* e = getenv("SPYB_TC_REWRITE_HI")) a.rewrite_hi = atoi(e) != 0;
    // cross terms as BF16 MMAs by default (1.1e-6 vs FP64, all-TF32: 1.3e-6); SPYB_TC_BF16=0 selects 3xTF32
    a.bf16_cross = 1;
    a.dbg = 0;
    if (const char* e = getenv("SPYB_TC_DBG")) a.dbg = atoi(e);
    if (const char* e = getenv("SPYB_TC_BF16")) a.bf16_cross = atoi(e) != 0;
    // 896 bytes of slack reach the next 1024-byte boundary from any 128-byte aligned start (dynamic shared memory
    // starts at least that aligned); 1024 would push the total 104 bytes past the 227 KB limit
    const size_t smem = 896 + (size_t)TC_STAGES * TC_STAGE + 16 * 8 + 8 * 32 * 16 * sizeof(float2) +
                        4 * 128 * sizeof(float);

    // per device / context attribute: set on every launch (several engines may live in one process)
    SPYB_CUDA(cudaFuncSetAttribute(csd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int dev = 0, n_sm = 148;
    SPYB_CUDA(cudaGetDevice(&dev));
    SPYB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    if (a.store_mode != 2 || a.n_freq_work <= 0 || a.n_freq_work > d.n_freq) a.n_freq_work = d.n_freq;
    const int n_items = a.store_mode == 3 ? d.n_freq : a.n_freq_work * a.n_tiles;
    const int grid = n_items < n_sm ? n_items : n_sm;
    csd_tc_kernel<<<grid, TC_THREADS, smem, stream>>>(tmap, a);
    SPYB_LAUNCH_CHECK("csd_tc_kernel");
    count_launch();
    return 0;
}

int csd_accumulate_tc(const CsdPlanarDesc& d, cudaStream_t stream) {
    if (reinterpret_cast<uintptr_t>(d.acc) % 16 != 0) return fail("tcgen05 CSD kernel needs 16-byte aligned buffers");
    TcArgs a = {};
    a.acc = reinterpret_cast<float2*>(d.acc);
    a.store_mode = 1;
    if (const char* e = getenv("SPYB_TC_STORE")) a.store_mode = atoi(e) != 0;
    return launch_tc(d, a, stream);
}

int csd_coherence_tc(const CsdPlanarDesc& d, int out_kind, void* out, cudaStream_t stream) {
    if (reinterpret_cast<uintptr_t>(out) % 16 != 0) return fail("tcgen05 CSD kernel needs 16-byte aligned buffers");
    TcArgs a = {};
    a.store_mode = 3;
    a.coh_out = out;
    a.out_kind = out_kind;
    return launch_tc(d, a, stream);
}

// Fused coherence of the frequencies this rank owns on several ranks: `d` covers the local slab, `slots` is the local
// slot buffer [n_src][d.n_freq][n_tiles][128][128] the peers filled (store_mode 2 with skip_own); source `skip_src`
// (this rank) is not read -- its contribution is the contraction of this launch.
int csd_coherence_tc_slots(const CsdPlanarDesc& d, const void* slots, int n_src, int skip_src, int out_kind, void* out,
                           cudaStream_t stream) {
    if (reinterpret_cast<uintptr_t>(out) % 16 != 0 || reinterpret_cast<uintptr_t>(slots) % 16 != 0)
        return fail("tcgen05 CSD kernel needs 16-byte aligned buffers");
    if (n_src < 1 || n_src > TC_MAX_RANKS) return fail("tile slots: 1..%d source ranks supported (got %d)", TC_MAX_RANKS, n_src);
    if (slots == nullptr) return fail("tile slots: no slot buffer");
    TcArgs a = {};
    a.store_mode = 3;
    a.coh_out = out;
    a.out_kind = out_kind;
    a.add_base = static_cast<const float2*>(slots);
    a.n_add_src = n_src;
    a.add_skip = skip_src;
    return launch_tc(d, a, stream);
}

int csd_tile_count(int n_chan) { return tri_tiles((n_chan + 127) / 128); }

int csd_accumulate_tc_tiles(const CsdPlanarDesc& d, void* const* owner_base, const int* f_begin, int n_owners,
                            int src_rank, int skip_own, cudaStream_t stream) {
    if (n_owners < 1 || n_owners > TC_MAX_RANKS) return fail("tile slots: 1..%d owner ranks supported (got %d)", TC_MAX_RANKS, n_owners);
    if (src_rank < 0) return fail("tile slots: bad source rank %d", src_rank);
    if (f_begin[0] != 0 || f_begin[n_owners] != d.n_freq) return fail("tile slots: frequency slabs must cover [0, %d)", d.n_freq);
    TcArgs a = {};
    a.store_mode = 2;
    a.n_owners = n_owners; a.src_rank = src_rank;
    for (int o = 0; o < n_owners; ++o) {
        if (f_begin[o + 1] < f_begin[o]) return fail("tile slots: frequency slabs must be ascending");
        if (owner_base[o] == nullptr || reinterpret_cast<uintptr_t>(owner_base[o]) % 16 != 0)
            return fail("tile slots: owner %d has no (aligned) slot buffer", o);
        a.owner_base[o] = reinterpret_cast<float2*>(owner_base[o]);
        a.f_begin[o] = f_begin[o];
    }
    a.f_begin[n_owners] = f_begin[n_owners];
    a.f_rot = n_owners > 1 ? f_begin[(src_rank + 1) % n_owners] % d.n_freq : 0;
    a.n_freq_work = d.n_freq;
    if (skip_own) {
        // the frequencies this rank owns are left out: it runs them itself in fused mode (csd_coherence_tc_slots)
        if (src_rank >= n_owners) return fail("tile slots: source rank %d owns no slab", src_rank);
        a.n_freq_work = d.n_freq - (f_begin[src_rank + 1] - f_begin[src_rank]);
        if (a.n_freq_work <= 0) return 0;
    }
    return launch_tc(d, a, stream);
}

}  // namespace spyb
