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
// One persistent CTA per SM; a work item is (frequency, 128-row block of the output):
//   warp 0      TMA producer: one 5-D tiled bulk-tensor load per stage brings KC rows of both planes
//               into shared memory, directly in the swizzled canonical MN-major layout (128B span, 32B atoms)
//   warps 2..5  converter: split the landed FP32 tile into hi (in place) and lo planes
//   warp 1      MMA issuer: 12 tcgen05.mma (M=128, N=C, K=8) per 8 rows, commits free the stage
//   warps 6..9  epilogue: tcgen05.ld -> alpha/beta -> global store, once per accumulation chain
// Accumulators: C_re in TMEM columns [0, C), C_im in [C, 2C); lane = output row.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "common.cuh"
#include "spyb_internal.h"

namespace spyb {
namespace {

constexpr int TC_KC = 16;              // rows per pipeline stage (two K=8 MMA steps)
constexpr int TC_THREADS = 320;        // 10 warps
constexpr int TC_CONV_THREADS = 128;   // warps 2..5
constexpr int TC_EPI_THREADS = 128;    // warps 6..9

struct TcArgs {
    int n_rows, n_freq, n_chan;
    int n_mblk;                // 128-row blocks of the output
    int cb_stride;             // 32-channel blocks per plane (= n_chan / 32)
    int n_stages;
    int tmem_cols;             // power of two >= 2*n_chan
    int chain_ksteps;          // pipeline stages (of TC_KC rows) accumulated in TMEM before a flush
    float alpha, beta;
    float2* acc;               // [n_freq][C][C]
};

// ---- PTX wrappers ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
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
// Instruction descriptor: TF32 x TF32 -> F32, both operands MN-major.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool neg_a) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((neg_a ? 1u : 0u) << 13) | (1u << 15) | (1u << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Work items: (frequency f, 128-row block mblk).  Only the upper block triangle is computed: block row
// mblk covers columns [128*mblk, C), so with two block rows the first is twice as heavy as the second.
// Round k of the persistent loop hands CTA b the unit u = b + k*grid; flipping the block row with the
// round parity (grid is even) makes every CTA alternate heavy / light items, while the two items of
// one frequency still run in the same round on neighbouring CTAs (their operand tile is shared in L2).
__device__ __forceinline__ void decode_item(int u, int round, int n_mblk, int& f, int& mblk) {
    f = u / n_mblk;
    mblk = u % n_mblk;
    if (n_mblk == 2) mblk ^= (round & 1);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
csd_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages][hi planes | lo planes] then barriers
    const uint32_t plane_bytes = (uint32_t)a.cb_stride * TC_KC * 128u;      // one plane of one stage
    const uint32_t half_bytes = 2u * plane_bytes;                          // re + im
    const uint32_t stage_bytes = 2u * half_bytes;                          // hi + lo
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)a.n_stages * stage_bytes);
    uint64_t* full_raw = bars;
    uint64_t* full_conv = bars + a.n_stages;
    uint64_t* empty = bars + 2 * a.n_stages;
    uint64_t* acc_full = bars + 3 * a.n_stages;
    uint64_t* acc_empty = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = a.n_chan;
    const int n_items = a.n_freq * a.n_mblk;
    const int n_ksteps = (a.n_rows + TC_KC - 1) / TC_KC;
    const int chain_ksteps = a.chain_ksteps;                               // stages per accumulation chain
    const uint32_t box_bytes = 4u * TC_KC * 128u;                          // bytes one TMA box delivers (128 channels)

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.n_stages; ++s) {
            mbar_init(&full_raw[s], 1);
            mbar_init(&full_conv[s], TC_CONV_THREADS);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, TC_EPI_THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            int s = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x, round = 0; item < n_items; item += gridDim.x, ++round) {
                int f, mblk;
                decode_item(item, round, a.n_mblk, f, mblk);
                // one box = 128 channels x KC rows of one plane; only channels >= 128*mblk are needed
                const int n_half = a.n_mblk - mblk;
                for (int ks = 0; ks < n_ksteps; ++ks) {
                    mbar_wait(&empty[s], phase ^ 1u);
                    mbar_arrive_expect_tx(&full_raw[s], box_bytes * 2u * (uint32_t)n_half);
                    uint8_t* dst = base + (size_t)s * stage_bytes;
                    for (int h = mblk; h < a.n_mblk; ++h) {
                        tma_load_5d(&tmap, &full_raw[s], dst + (size_t)h * box_bytes, 0, ks * TC_KC, 4 * h, 0, f);
                        tma_load_5d(&tmap, &full_raw[s], dst + plane_bytes + (size_t)h * box_bytes, 0, ks * TC_KC, 4 * h, 1, f);
                    }
                    if (++s == a.n_stages) { s = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t lbo = TC_KC * 128u, sbo = 512u;
        int s = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int item = blockIdx.x, round = 0; item < n_items; item += gridDim.x, ++round) {
            int f, mblk;
            decode_item(item, round, a.n_mblk, f, mblk);
            const int N = C - 128 * mblk;                          // columns [128*mblk, C)
            const uint32_t idesc_pos = make_idesc(128, N, false);
            const uint32_t idesc_neg = make_idesc(128, N, true);
            const uint32_t d_re = tmem_base, d_im = tmem_base + (uint32_t)N;
            for (int ks = 0; ks < n_ksteps; ++ks) {
                const int kc = ks % chain_ksteps;                  // position inside the accumulation chain
                if (kc == 0) {                                     // the epilogue must have drained the previous chain
                    mbar_wait(acc_empty, acc_phase ^ 1u);
                    tc_fence_after();
                }
                mbar_wait(&full_conv[s], phase);
                tc_fence_after();
                const bool chain_end = (kc == chain_ksteps - 1) || (ks == n_ksteps - 1);
                if (elect_one()) {
                    const uint32_t st = smem_u32(base + (size_t)s * stage_bytes);
                    const uint32_t a_off = (uint32_t)mblk * 4u * lbo;      // 128 channels = 4 atoms of 32
#pragma unroll
                    for (int kk = 0; kk < TC_KC / 8; ++kk) {
                        const uint32_t k_off = (uint32_t)kk * 1024u;
                        // rows and columns of this item both start at channel 128*mblk: A and B share descriptors
                        const uint32_t re_hi = st + k_off + a_off, im_hi = st + plane_bytes + k_off + a_off;
                        const uint32_t re_lo = re_hi + half_bytes, im_lo = im_hi + half_bytes;
                        const uint64_t Dre_hi = make_smem_desc(re_hi, lbo, sbo), Dre_lo = make_smem_desc(re_lo, lbo, sbo);
                        const uint64_t Dim_hi = make_smem_desc(im_hi, lbo, sbo), Dim_lo = make_smem_desc(im_lo, lbo, sbo);
                        const uint32_t first = (kc > 0 || kk > 0) ? 1u : 0u;
                        // C_re = Re^T Re + Im^T Im   (A = first operand, B = second; both read the same tile)
                        umma_tf32(d_re, Dre_lo, Dre_hi, idesc_pos, first);
                        umma_tf32(d_re, Dre_hi, Dre_lo, idesc_pos, 1u);
                        umma_tf32(d_re, Dim_lo, Dim_hi, idesc_pos, 1u);
                        umma_tf32(d_re, Dim_hi, Dim_lo, idesc_pos, 1u);
                        umma_tf32(d_re, Dre_hi, Dre_hi, idesc_pos, 1u);
                        umma_tf32(d_re, Dim_hi, Dim_hi, idesc_pos, 1u);
                        // C_im = Im^T Re - Re^T Im
                        umma_tf32(d_im, Dim_lo, Dre_hi, idesc_pos, first);
                        umma_tf32(d_im, Dim_hi, Dre_lo, idesc_pos, 1u);
                        umma_tf32(d_im, Dre_lo, Dim_hi, idesc_neg, 1u);
                        umma_tf32(d_im, Dre_hi, Dim_lo, idesc_neg, 1u);
                        umma_tf32(d_im, Dim_hi, Dre_hi, idesc_pos, 1u);
                        umma_tf32(d_im, Dre_hi, Dim_hi, idesc_neg, 1u);
                    }
                    umma_commit(&empty[s]);                      // stage free once these MMAs retire
                    if (chain_end) umma_commit(acc_full);
                }
                __syncwarp();
                if (chain_end) acc_phase ^= 1u;
                if (++s == a.n_stages) { s = 0; phase ^= 1u; }
            }
        }
    } else if (warp < 6) {
        // ===== converter (warps 2..5): FP32 tile -> hi (in place) + lo =====
        const int ct = threadIdx.x - 64;                        // 0..127
        int s = 0;
        uint32_t phase = 0;
        for (int item = blockIdx.x, round = 0; item < n_items; item += gridDim.x, ++round) {
            int f, mblk;
            decode_item(item, round, a.n_mblk, f, mblk);
            const uint32_t skip_chunks = (uint32_t)mblk * (box_bytes / 16u);          // channels below 128*mblk are not loaded
            const uint32_t plane_item_chunks = (uint32_t)(a.n_mblk - mblk) * (box_bytes / 16u);
            const uint32_t item_chunks = 2u * plane_item_chunks;
            for (int ks = 0; ks < n_ksteps; ++ks) {
                mbar_wait(&full_raw[s], phase);
                float4* hi = reinterpret_cast<float4*>(base + (size_t)s * stage_bytes);
                float4* lo = reinterpret_cast<float4*>(base + (size_t)s * stage_bytes + half_bytes);
                for (uint32_t q0 = ct; q0 < item_chunks; q0 += TC_CONV_THREADS) {
                    // 16-byte chunk q0 of the loaded part -> (plane, offset inside the plane)
                    const uint32_t pl = q0 / plane_item_chunks, rem = q0 - pl * plane_item_chunks;
                    const uint32_t q = pl * (plane_bytes / 16u) + skip_chunks + rem;
                    const float4 x = hi[q];
                    float4 h, l;
                    h.x = rna_tf32(x.x); h.y = rna_tf32(x.y); h.z = rna_tf32(x.z); h.w = rna_tf32(x.w);
                    l.x = rna_tf32(x.x - h.x); l.y = rna_tf32(x.y - h.y);
                    l.z = rna_tf32(x.z - h.z); l.w = rna_tf32(x.w - h.w);
                    hi[q] = h;
                    lo[q] = l;
                }
                fence_proxy_async_smem();
                mbar_arrive(&full_conv[s]);
                if (++s == a.n_stages) { s = 0; phase ^= 1u; }
            }
        }
    } else {
        // ===== epilogue (warps 6..9): TMEM -> registers -> global, once per accumulation chain =====
        // The tensor core adds into its FP32 accumulators with truncation, so the error of a chain grows
        // linearly with the number of MMAs feeding one accumulator (~3e-8 per MMA, measured).  Chains are
        // therefore cut after `chain_ksteps` stages and combined here with ordinary round-to-nearest FP32
        // adds; the partial tile a later chain re-reads was written by this CTA moments ago (L2 hits).
        const int lane_grp = warp & 3;                          // TMEM lanes this warp may read
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x, round = 0; item < n_items; item += gridDim.x, ++round) {
            int f, mblk;
            decode_item(item, round, a.n_mblk, f, mblk);
            // Thread = output row i, columns n0 + [0, N).  Elements above the diagonal are stored twice (as is,
            // and conjugated into the mirrored position -- lanes hold consecutive i, so those stores coalesce);
            // nothing below the diagonal is used, which makes the result exactly Hermitian.
            const int n0 = 128 * mblk, N = C - n0;
            const int i = n0 + lane_grp * 32 + lane;
            float2* __restrict__ fmat = a.acc + (size_t)f * C * C;
            float2* __restrict__ orow = fmat + (size_t)i * C;
            const uint32_t t_row = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
            const int i_warp_min = n0 + lane_grp * 32;             // smallest row of this warp
            for (int k0 = 0; k0 < n_ksteps; k0 += chain_ksteps) {
                const float beta = k0 == 0 ? a.beta : 1.f;
                mbar_wait(acc_full, acc_phase);
                tc_fence_after();
                for (int c0 = 0; c0 < N; c0 += 16) {
                    const int j0 = n0 + c0;
                    if (j0 + 15 < i_warp_min) continue;            // chunk entirely below the diagonal (warp-uniform)
                    uint32_t vr[16], vi[16];
                    tmem_ld16(t_row + (uint32_t)c0, vr);
                    tmem_ld16(t_row + (uint32_t)(N + c0), vi);
                    tmem_ld_wait();
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) {
                        const int j = j0 + jj;
                        if (j < i) continue;
                        float2 o = make_float2(__uint_as_float(vr[jj]) * a.alpha, __uint_as_float(vi[jj]) * a.alpha);
                        float2* d0 = orow + j;
                        if (beta != 0.f) {
                            const float2 old0 = *d0;
                            o.x += beta * old0.x; o.y += beta * old0.y;
                        }
                        if (j == i) o.y = 0.f;                      // auto-spectra are exactly real
                        *d0 = o;
                        if (j != i) fmat[(size_t)j * C + i] = make_float2(o.x, -o.y);
                    }
                }
                tc_fence_before();
                mbar_arrive(acc_empty);
                acc_phase ^= 1u;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
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
    return (n_chan == 128 || n_chan == 256) && sx_r % 4 == 0 && sx_f % 4 == 0;
}

int csd_accumulate_tc(const CsdPlanarDesc& d, cudaStream_t stream) {
    if (d.n_freq <= 0 || d.n_chan <= 0) return 0;
    if (!csd_tc_supported(d.n_chan, d.sx_f, d.sx_r))
        return fail("tcgen05 CSD kernel needs n_chan in {128, 256} and 16-byte aligned strides "
                    "(got n_chan=%d, sx_f=%lld, sx_r=%lld)", d.n_chan, d.sx_f, d.sx_r);
    if (reinterpret_cast<uintptr_t>(d.planes) % 16 != 0 || reinterpret_cast<uintptr_t>(d.acc) % 16 != 0)
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

    TcArgs a;
    a.n_rows = d.n_rows; a.n_freq = d.n_freq; a.n_chan = C;
    a.n_mblk = (C + 127) / 128;
    a.cb_stride = C / 32;
    a.alpha = d.alpha; a.beta = d.beta;
    a.acc = reinterpret_cast<float2*>(d.acc);
    a.tmem_cols = 32;
    while (a.tmem_cols < 2 * C) a.tmem_cols <<= 1;
    // accumulation-chain length in rows (multiple of TC_KC); SPYB_TC_CHAIN_ROWS overrides for experiments
    int chain_rows = 128;
    if (const char* e = getenv("SPYB_TC_CHAIN_ROWS")) chain_rows = atoi(e);
    if (chain_rows < TC_KC) chain_rows = TC_KC;
    a.chain_ksteps = chain_rows / TC_KC;
    const size_t stage_bytes = (size_t)4 * a.cb_stride * TC_KC * 128;
    a.n_stages = (int)((200 * 1024) / stage_bytes);
    if (a.n_stages > 8) a.n_stages = 8;
    if (a.n_stages < 2) return fail("csd_tc: stage does not fit shared memory");
    const size_t smem = 1024 + (size_t)a.n_stages * stage_bytes + (3 * a.n_stages + 2) * 8 + 16;

    static bool configured = false;
    if (!configured) {
        SPYB_CUDA(cudaFuncSetAttribute(csd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    int dev = 0, n_sm = 148;
    SPYB_CUDA(cudaGetDevice(&dev));
    SPYB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    const int n_items = d.n_freq * a.n_mblk;
    int grid = n_items < n_sm ? n_items : n_sm;
    if (a.n_mblk == 2) grid &= ~1;      // decode_item() pairs the two block rows of a frequency within a round
    csd_tc_kernel<<<grid, TC_THREADS, smem, stream>>>(tmap, a);
    SPYB_LAUNCH_CHECK("csd_tc_kernel");
    count_launch();
    return 0;
}

}  // namespace spyb
