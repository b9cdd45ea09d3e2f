// K1, four-step variant for the headline launch shape: N = 4096 samples, one taper, planar result (what the
// cross-spectral kernel consumes), channel counts in multiples of 32.  EXPERIMENTAL, opt-in (SPYB_MTM_4S=1): measured
// on B200 at cfg-2 0.567 ms without de-meaning and 0.73 ms with it, against 0.65 ms of mtm_tma.cu.
//
// Why: mtm_tma.cu transforms 8 channels x 4096 samples per block in shared memory, so everything it moves through the
// memory system is a 32-byte piece -- 32-byte rows of the trial on the way in, 32-byte pieces of the planar result
// 400 KB apart on the way out -- and the kernel sits at 2.5 TB/s, the rate at which B200 handles 32-byte sectors
// (tools/micro/scatter_store3.cu: the same bytes leave in 0.32 ms as 32-byte pieces, in 0.17 ms as 128-byte lines).
// Here a work unit is 32 channels (16 complex pairs = one 128-byte line per sample) of one trial, and the 4096-point
// transform is split 64 x 64 (n = 64 n1 + n2, k = k1 + 64 k2):
//   phase A (item = unit x SUB values of n2): 64-point DFTs over n1 on rows that are 128 bytes wide, twiddle
//            W_4096^(n2 k1), result Y[k1][n2][pair] into a ring buffer that stays in L2 (ncu: 0.90 GB of DRAM
//            writes per launch for 0.84 GB of results);
//   phase B (item = unit x SUB values of k1, closed under k1 -> 64 - k1): 64-point DFTs over n2, the pair split
//            X_c = (Z[k] + conj Z[N-k]) / 2, X_{c+1} = (Z[k] - conj Z[N-k]) / 2i, scale, and the planar result written
//            as full 128-byte lines.
// Both phases run in ONE persistent kernel (four 128-thread blocks per SM).  Items are numbered globally
// [A0] [A1 B0] [A2 B1] ... and item g belongs to block g mod grid; a phase-B item waits on a per-unit counter that the
// phase-A items of its unit bump after their stores (release / acquire through __threadfence and an atomic), a
// phase-A item that reuses a ring slot waits for the phase-B items that read it.  Every wait refers to items that come
// earlier in the global order, and an item never waits on behalf of a later one (the next item's input is fetched one
// item ahead only if a LOOK at its counter says it is there), so with all blocks resident -- grid = what the
// occupancy calculator allows, nothing else running on the device -- the earliest unfinished item can always run.
// A kernel that shares the device with other work would need a ticket counter instead of the static assignment.
//
// Every 64-point DFT is two radix-8 butterflies in registers with one exchange through shared memory; a thread
// carries two complex pairs (16 bytes), a warp instruction touches 512 contiguous bytes of shared or global memory.
//
// De-meaning (polyremoval = 0, scipy.signal.detrend(type='constant'), compRoutines.py:169-172) needs the mean of all
// 4096 samples before the taper is applied, which no phase-A item sees.  Every item subtracts the channel's FIRST
// sample x0 instead (exact, whatever the offset of the recording) and writes the sum of (x - x0) over its samples; the
// phase-A item of a unit that finishes last forms delta = mean - x0 (fixed summation order), and phase B removes
// delta * [phase A applied to the taper] from the rows it loads -- there the term is 64 samples deep; removing
// delta * DFT(taper) from the finished spectrum instead loses 1.7e-5 in the bins 0 and +-1 to cancellation.
//
// Same arithmetic as mtm_tma.cu / mtm_dif.cu otherwise; replaces syncopy/specest/mtmfft.py:111-127.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "mtm_args.cuh"
#include "packed.cuh"
#include "spyb_internal.h"

namespace spyb {
namespace {

constexpr int N = 4096;
constexpr int SUB = 2;                             // values of n2 (phase A) / k1 (phase B) per item
constexpr int SUBBITS = 1;
constexpr int THREADS = 64 * SUB;
constexpr int NF = N / 2 + 1;
constexpr int ITEMS = 64 / SUB;                    // items per unit and phase
constexpr int Y_UNIT_ELEMS = 64 * 64 * 8;          // 16-byte elements (two pairs each) per unit = 512 KB
constexpr int EX_ELEMS = 8 * 8 * SUB * 8;          // exchange buffer, 16-byte elements (16 KB); Z buffer: the same size
constexpr int NWARPS = THREADS / 32;
constexpr int SMEM_BYTES = 2 * EX_ELEMS * 16 + 64 * 8 + NWARPS * 8 * 16 + 16;     // last 16 bytes: a flag

struct FsArgs {
    const float* x;
    long long trial_stride;
    int n_chan, groups, n_units;
    const float* taper;        // [4096]
    float half_scale;
    int detrend;               // 0: none, 1: remove the mean
    float* out;
    long long so_trial, so_freq;
    ulonglong2* Y;             // ring [ring_units][64 k1][64 n2][8 x 16 B]
    int ring_units, ub, n_batches;
    int* cnt_m;                // [n_units] 1 once the unit's delta is written
    int* cnt_a;                // [n_units] phase-A items finished
    int* cnt_b;                // [n_units] phase-B items finished
    float4* psum;              // [n_units][ITEMS][8] partial sums of (x - x0), 4 channels each
    float4* delta;             // [n_units][8] mean - x0
    const float2* what2;       // [64 k1][64 n2] phase A applied to the taper
    const float2* tw4096;      // [4096] W_4096^m
    int dbg;
};

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Bounded spin on a counter another block bumps: a scheduling bug must end in a launch error, not a hung GPU.
__device__ __forceinline__ void wait_count(const int* p, int target) {
    unsigned long long t0 = 0;
    for (unsigned spin = 0;; ++spin) {
        int v;
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        if (v >= target) return;
        __nanosleep(32);
        if ((spin & 1023u) == 1023u) {
            const unsigned long long now = gtimer();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) __trap();
        }
    }
}
// release: everything this block wrote before the preceding __syncthreads becomes visible before the counter moves
__device__ __forceinline__ void bump(int* p) {
    __threadfence();
    atomicAdd(p, 1);
}

// k1 values of phase-B item i, slot l (0..SUB-1): pairs {k1, 64 - k1}; item 0 holds the self-paired 0 and 32
__device__ __forceinline__ int k1_of(int i, int l) {
    const int base = (SUB / 2) * i + (l >> 1), odd = l & 1;
    if (base == 0) return odd ? 32 : 0;
    return odd ? 64 - base : base;
}

struct Item { int kind, unit, sub; };              // kind: 0 phase A, 1 phase B, -1 none

template <int MINB>
__global__ void __launch_bounds__(THREADS, MINB) mtm_4s_kernel(const FsArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    ulonglong2* ex = reinterpret_cast<ulonglong2*>(smem);                       // [8 q][8 j][SUB][8 pp]
    ulonglong2* zb = ex + EX_ELEMS;                                             // [SUB][64 k2][8 pp]
    float2* w64 = reinterpret_cast<float2*>(smem + 2 * EX_ELEMS * 16);          // W_64^m
    float4* red = reinterpret_cast<float4*>(smem + 2 * EX_ELEMS * 16 + 64 * 8); // [warps][8]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pp = tid & 7, mid = (tid >> 3) & (SUB - 1), top = tid >> (3 + SUBBITS);       // top: 0..7
    // phase B, first half: slot l and row residue j of this thread (the low bits of j next to pp: a warp reads whole lines)
    const int bj = mid + SUB * (top % (8 / SUB)), bl = top / (8 / SUB);
    if (tid < 64) w64[tid] = a.tw4096[tid * 64];
    __syncthreads();
    const c2 hs = bc(a.half_scale);

    // global item order: [A0] [A1 B0] [A2 B1] ... [B_last]; item g -> block g % gridDim.x.  Every wait of an item
    // refers to items that come earlier in this order (or, for the partial sums, to the publishing step at the very
    // beginning of items of the same batch, which waits for nothing).
    const int S = a.ub * ITEMS;
    auto decode = [&](long long g) {
        Item it; it.kind = -1; it.unit = 0; it.sub = 0;
        int round, o;
        if (g < S) { round = 0; o = (int)g; }
        else { const long long h = g - S; round = 1 + (int)(h / (2 * S)); o = (int)(h % (2 * S)); }
        if (round > a.n_batches) { it.kind = -2; return it; }                 // past the end
        if (round == 0 || o < S) {
            if (round == a.n_batches) return it;
            it.unit = round * a.ub + o / ITEMS; it.sub = o % ITEMS;
            it.kind = it.unit < a.n_units ? 0 : -1;
        } else {
            o -= S;
            it.unit = (round - 1) * a.ub + o / ITEMS; it.sub = o % ITEMS;
            it.kind = it.unit < a.n_units ? 1 : -1;
        }
        return it;
    };
    auto next_valid = [&](long long& g) {                                      // first real item at or after g
        for (;; g += gridDim.x) {
            const Item it = decode(g);
            if (it.kind != -1) return it;
        }
    };

    float4 pre[8];                                 // input of the item that comes next, loaded one item ahead
    float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f);   // phase A with de-meaning: first sample of the trial (this thread's channels)
    int* flag = reinterpret_cast<int*>(smem + SMEM_BYTES - 16);
    auto issue_loads = [&](const Item& it) {
        if (it.kind == 0) {
            // thread (j = top, n2l = mid, pp): samples n = 64 (j + 8 r) + SUB b + n2l, channels 32 g + 4 pp .. + 3
            const int trial = it.unit / a.groups, g = it.unit - trial * a.groups;
            const float* __restrict__ x00 = a.x + (long long)trial * a.trial_stride + 32 * g + 4 * pp;
            const float* __restrict__ xp = x00 + (long long)(64 * top + SUB * it.sub + mid) * a.n_chan;
#pragma unroll
            for (int r = 0; r < 8; ++r) pre[r] = __ldcs(reinterpret_cast<const float4*>(xp + (long long)(512 * r) * a.n_chan));
            if (a.detrend && !(a.dbg & 16)) x0 = __ldg(reinterpret_cast<const float4*>(x00));
        } else {
            // thread (l = bl, j = bj, pp): rows n2 = j + 8 r of Y[k1]
            const ulonglong2* __restrict__ yu = a.Y + (long long)(it.unit % a.ring_units) * Y_UNIT_ELEMS;
            const int k1 = k1_of(it.sub, bl), j = bj;
#pragma unroll
            for (int r = 0; r < 8; ++r) pre[r] = __ldcg(reinterpret_cast<const float4*>(yu + (k1 * 64 + j + 8 * r) * 8 + pp));
        }
    };

    // warp 0: the unit whose phase-A item this block released last, and the counter value the release returned
    int pend_unit = -1, pend_old = 0;
    auto resolve_pending = [&]() {              // warp 0 only
        if (pend_unit < 0) return;
        const int last = __shfl_sync(0xffffffffu, pend_old == ITEMS - 1, 0);
        if (last) {
            // this block's item was the last of its unit: delta = mean - x0 from the ITEMS partial sums (fixed order)
            __threadfence();
            const int part = lane >> 3;                                        // 4 parts x ITEMS / 4 partial sums each
            const float4* __restrict__ ps = a.psum + ((long long)pend_unit * ITEMS + part * (ITEMS / 4)) * 8 + (lane & 7);
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int b = 0; b < ITEMS / 4; ++b) {
                const float4 p = __ldcg(ps + b * 8);
                s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
            }
            s.x += __shfl_xor_sync(0xffffffffu, s.x, 8);  s.y += __shfl_xor_sync(0xffffffffu, s.y, 8);
            s.z += __shfl_xor_sync(0xffffffffu, s.z, 8);  s.w += __shfl_xor_sync(0xffffffffu, s.w, 8);
            s.x += __shfl_xor_sync(0xffffffffu, s.x, 16); s.y += __shfl_xor_sync(0xffffffffu, s.y, 16);
            s.z += __shfl_xor_sync(0xffffffffu, s.z, 16); s.w += __shfl_xor_sync(0xffffffffu, s.w, 16);
            const float inv = 1.f / (float)N;
            if (lane < 8) a.delta[(long long)pend_unit * 8 + lane] = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
            __syncwarp();
            if (lane == 0) bump(a.cnt_m + pend_unit);                          // delta of the unit is there
        }
        pend_unit = -1;
    };
    // phase-B items need every phase-A item of their unit and, with de-meaning, the unit's delta
    auto wait_producers = [&](int unit) {       // thread 0
        wait_count(a.cnt_a + unit, ITEMS);
        if (a.detrend && !(a.dbg & 2)) wait_count(a.cnt_m + unit, 1);
    };

    long long g = blockIdx.x;
    Item cur = next_valid(g);
    if (cur.kind >= 0) {
        if (cur.kind == 1) { if (tid == 0) wait_producers(cur.unit); __syncthreads(); }
        issue_loads(cur);
    }
    while (cur.kind >= 0) {
        long long gn = g + gridDim.x;
        const Item nxt = next_valid(gn);
        const int trial = cur.unit / a.groups, grp = cur.unit - trial * a.groups;
        c2 va[8], vb[8];
        if (cur.kind == 0) {
            // ---------------- phase A ----------------
            const int j = top, n2 = SUB * cur.sub + mid;
            if (cur.unit >= a.ring_units) {        // the ring slot was read by the phase-B items of an earlier unit
                if (tid == 0) wait_count(a.cnt_b + (cur.unit - a.ring_units), ITEMS);
                __syncthreads();
            }
            float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float w = __ldg(a.taper + 64 * (j + 8 * r) + n2);
                const float d0 = pre[r].x - x0.x, d1 = pre[r].y - x0.y, d2 = pre[r].z - x0.z, d3 = pre[r].w - x0.w;
                s4.x += d0; s4.y += d1; s4.z += d2; s4.w += d3;
                va[r] = mul2(pk(d0, d1), bc(w));
                vb[r] = mul2(pk(d2, d3), bc(w));
            }
            if (a.detrend && !(a.dbg & 8)) {
                s4.x += __shfl_xor_sync(0xffffffffu, s4.x, 8);  s4.y += __shfl_xor_sync(0xffffffffu, s4.y, 8);
                s4.z += __shfl_xor_sync(0xffffffffu, s4.z, 8);  s4.w += __shfl_xor_sync(0xffffffffu, s4.w, 8);
                s4.x += __shfl_xor_sync(0xffffffffu, s4.x, 16); s4.y += __shfl_xor_sync(0xffffffffu, s4.y, 16);
                s4.z += __shfl_xor_sync(0xffffffffu, s4.z, 16); s4.w += __shfl_xor_sync(0xffffffffu, s4.w, 16);
                if (lane < 8) red[warp * 8 + lane] = s4;
            }
            dft8(va);
            dft8(vb);
            // T[j][q] = W_64^(j q) * sum_r v[j + 8 r] W_8^(r q)  ->  exchange [q][j][n2l][pp]
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                c2 pa = va[q], pb = vb[q];
                if (q > 0) {
                    const float2 w = w64[j * q];
                    pa = cmul2(pa, w.x, w.y);
                    pb = cmul2(pb, w.x, w.y);
                }
                ex[((q * 8 + j) * SUB + mid) * 8 + pp] = make_ulonglong2(pa, pb);
            }
        } else {
            // ---------------- phase B, first half ----------------
            const int j = bj, l = bl;
#pragma unroll
            for (int r = 0; r < 8; ++r) { va[r] = pk(pre[r].x, pre[r].y); vb[r] = pk(pre[r].z, pre[r].w); }
            if (a.detrend && !(a.dbg & 4)) {
                // the samples went in as (x - x0); remove (mean - x0) * [phase A applied to the taper] here, where the
                // term is still small (64 samples deep) instead of in the spectrum's bins 0 and +-1
                const float4 dl = __ldcg(a.delta + (long long)cur.unit * 8 + pp);
                const c2 da = pk(dl.x, dl.y), db = pk(dl.z, dl.w);
                const float2* __restrict__ wt = a.what2 + k1_of(cur.sub, l) * 64 + j;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float2 w = __ldg(wt + 8 * r);
                    va[r] = sub2(va[r], cmul2(da, w.x, w.y));
                    vb[r] = sub2(vb[r], cmul2(db, w.x, w.y));
                }
            }
            dft8(va);
            dft8(vb);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                c2 pa = va[q], pb = vb[q];
                if (q > 0) {
                    const float2 w = w64[j * q];
                    pa = cmul2(pa, w.x, w.y);
                    pb = cmul2(pb, w.x, w.y);
                }
                ex[((q * 8 + j) * SUB + l) * 8 + pp] = make_ulonglong2(pa, pb);
            }
        }
        // The next item's input is fetched one item ahead when its producers (phase-A items of other blocks) are
        // already through.  Only a look, never a wait: this item must finish whatever the state of later ones.
        if (tid == 0) {
            int ready = 1;
            if (nxt.kind == 1) {
                int v;
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(a.cnt_a + nxt.unit) : "memory");
                ready = v >= ITEMS;
                if (ready && a.detrend && !(a.dbg & 2)) {
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(a.cnt_m + nxt.unit) : "memory");
                    ready = v >= 1;
                }
            }
            *flag = ready;
        }
        __syncthreads();
        const bool fetched = nxt.kind >= 0 && *flag != 0;
        if (warp == 0) resolve_pending();      // the release of the previous item has long returned by now
        // phase B: every thread has consumed its rows of the ring slot -> the slot may be overwritten
        if (cur.kind == 1 && tid == 0) atomicAdd(a.cnt_b + cur.unit, 1);
        if (cur.kind == 0 && a.detrend && !(a.dbg & 8) && tid < 8) {
            float4 s = red[tid];
#pragma unroll
            for (int w = 1; w < NWARPS; ++w) {      // fixed order: deterministic
                const float4 p = red[w * 8 + tid];
                s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
            }
            a.psum[((long long)cur.unit * ITEMS + cur.sub) * 8 + tid] = s;
        }
        if (fetched) issue_loads(nxt);
        {
            const int q = top;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const ulonglong2 v = ex[((q * 8 + jj) * SUB + mid) * 8 + pp];
                va[jj] = v.x; vb[jj] = v.y;
            }
            dft8(va);
            dft8(vb);
        }
        if (cur.kind == 0) {
            // thread (q = top, n2l = mid, pp): Y[q + 8 c][n2] = W_4096^(n2 k1) * sum_j T[j][q] W_8^(j c)
            const int q = top, n2 = SUB * cur.sub + mid;
            ulonglong2* __restrict__ yu = a.Y + (long long)(cur.unit % a.ring_units) * Y_UNIT_ELEMS;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int k1 = q + 8 * c;
                c2 pa = va[c], pb = vb[c];
                const float2 w = __ldg(a.tw4096 + n2 * k1);
                pa = cmul2(pa, w.x, w.y);
                pb = cmul2(pb, w.x, w.y);
                yu[(k1 * 64 + n2) * 8 + pp] = make_ulonglong2(pa, pb);
            }
            __syncthreads();                   // all stores issued; the exchange buffer is free again
            if (warp == 0) {
                // Release the item.  With de-meaning the counter's old value tells which item of the unit came last; the
                // answer is only looked at one item later (resolve_pending), so nobody waits for the atomic's round trip.
                if (lane == 0) {
                    __threadfence();
                    if (a.detrend && !(a.dbg & 2)) pend_old = atomicAdd(a.cnt_a + cur.unit, 1);
                    else atomicAdd(a.cnt_a + cur.unit, 1);
                }
                if (a.detrend && !(a.dbg & 2)) pend_unit = cur.unit;
            }
        } else {
            // thread (q = top, l = mid, pp): Z[k1 + 64 (q + 8 c)] -> Z[l][k2][pp] (its own buffer: the previous item's
            // readers of it are past the barrier above)
            const int q = top, l = mid;
#pragma unroll
            for (int c = 0; c < 8; ++c) zb[(l * 64 + q + 8 * c) * 8 + pp] = make_ulonglong2(va[c], vb[c]);
            __syncthreads();                   // also: everyone is done with the exchange buffer
            // output: SUB slots x 32 values of k2 (k <= 2047) = 32 SUB bins, four per (tid >> 3)
            float* __restrict__ op0 = a.out + (long long)trial * a.so_trial + 32 * grp + 4 * pp;
            auto emit = [&](int lo, int k2) {
                const int k1 = k1_of(cur.sub, lo), k = k1 + 64 * k2;
                const int kp = (N - k) & (N - 1);
                const int lp = (cur.sub == 0 && lo < 2) ? lo : (lo ^ 1);
                const ulonglong2 z1 = zb[(lo * 64 + k2) * 8 + pp];
                const ulonglong2 z2 = zb[(lp * 64 + (kp >> 6)) * 8 + pp];
                const c2 s0 = add2(z1.x, z2.x), d0 = sub2(z1.x, z2.x), s1 = add2(z1.y, z2.y), d1 = sub2(z1.y, z2.y);
                const c2 re0 = mul2(s0, hs), im0 = mul2(mul_mi(d0), hs), re1 = mul2(s1, hs), im1 = mul2(mul_mi(d1), hs);
                const float4 vr = make_float4(re(re0), im(re0), re(re1), im(re1));
                const float4 vi = make_float4(re(im0), im(im0), re(im1), im(im1));
                if ((a.dbg & 1) && vr.x != 12345.678f) return;
                float* op = op0 + (long long)k * a.so_freq;
                __stcs(reinterpret_cast<float4*>(op), vr);
                __stcs(reinterpret_cast<float4*>(op + a.n_chan), vi);
            };
#pragma unroll
            for (int rd = 0; rd < 4; ++rd) {
                const int idx = (tid >> 3) + 8 * SUB * rd;
                emit(idx >> 5, idx & 31);
            }
            if (cur.sub == 0 && tid < 8) emit(0, 32);   // Nyquist bin k = 2048 = 0 + 64 * 32, its own partner
        }
        if (nxt.kind >= 0 && !fetched) {       // its producers were not through yet: wait for them now, nothing is held up
            if (tid == 0) wait_producers(nxt.unit);
            __syncthreads();
            issue_loads(nxt);
        }
        cur = nxt;
        g = gn;
    }
    if (warp == 0) resolve_pending();
}

// Phase A applied to the taper itself: what2[k1][n2] = W_4096^(n2 k1) * sum_n1 w[64 n1 + n2] W_64^(n1 k1), float64 sums
__global__ void __launch_bounds__(256) taper_phase_a_kernel(const float* __restrict__ w, float2* __restrict__ what2) {
    const int idx = blockIdx.x * 256 + threadIdx.x;          // k1 * 64 + n2
    if (idx >= N) return;
    const int k1 = idx >> 6, n2 = idx & 63;
    double sr = 0.0, si = 0.0;
    for (int n1 = 0; n1 < 64; ++n1) {
        const int p = (64 * n1 * k1 + n2 * k1) & (N - 1);   // phase of W_4096^((64 n1 + n2) k1)
        double sn, cs;
        sincospi(-2.0 * (double)p / (double)N, &sn, &cs);
        const double wv = (double)w[64 * n1 + n2];
        sr += wv * cs; si += wv * sn;
    }
    what2[idx] = make_float2((float)sr, (float)si);
}

// ---- per-device workspace: ring buffer, counters, partial sums, tables --------------------------------------------
struct Workspace {
    int device = -1;
    cudaStream_t stream = nullptr;
    ulonglong2* Y = nullptr;
    int ring_units = 0;
    int* cnt = nullptr;        // [3][cap_units]
    float4* psum = nullptr;
    float4* delta = nullptr;
    float2* what2 = nullptr;
    int cap_units = 0;
    float2* tw4096 = nullptr;
    int grid = 0, ub = 0;
};
constexpr int MINB = 4;
const auto KERNEL = mtm_4s_kernel<MINB>;
std::mutex g_mu;
std::vector<Workspace> g_ws;

Workspace* get_workspace(int dev, cudaStream_t stream, int n_units) {
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace* ws = nullptr;
    for (auto& w : g_ws)
        if (w.device == dev && w.stream == stream) ws = &w;
    if (!ws) {
        if (g_ws.size() >= 8) return nullptr;          // one workspace per (device, stream); more streams -> other kernel
        g_ws.reserve(8);
        g_ws.emplace_back();
        ws = &g_ws.back();
        ws->device = dev; ws->stream = stream;
        int n_sm = 148, per_sm = 0;
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return nullptr;
        if (cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) return nullptr;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, KERNEL, THREADS, SMEM_BYTES) != cudaSuccess || per_sm < 1)
            return nullptr;
        if (per_sm > MINB) per_sm = MINB;
        ws->grid = n_sm * per_sm;
        ws->ub = (ws->grid + ITEMS - 1) / ITEMS;       // about one phase-A item per block and round
        ws->ring_units = 3 * ws->ub;
        std::vector<float2> tw(N);
        for (int m = 0; m < N; ++m) {
            const double ang = -2.0 * 3.14159265358979323846 * (double)m / (double)N;
            tw[m] = make_float2((float)std::cos(ang), (float)std::sin(ang));
        }
        if (cudaMalloc(&ws->Y, (size_t)ws->ring_units * Y_UNIT_ELEMS * 16) != cudaSuccess) return nullptr;
        if (cudaMalloc(&ws->tw4096, N * sizeof(float2)) != cudaSuccess) return nullptr;
        if (cudaMalloc(&ws->what2, N * sizeof(float2)) != cudaSuccess) return nullptr;
        if (cudaMemcpy(ws->tw4096, tw.data(), N * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    }
    if (n_units > ws->cap_units) {
        // grown while earlier launches on this stream may still run: stream-ordered free, plain allocation
        if (ws->cnt) {
            cudaStreamSynchronize(stream);
            cudaFree(ws->cnt); cudaFree(ws->psum); cudaFree(ws->delta);
            ws->cnt = nullptr; ws->psum = nullptr; ws->delta = nullptr;
        }
        const int cap = n_units + n_units / 4;
        if (cudaMalloc(&ws->cnt, (size_t)3 * cap * sizeof(int)) != cudaSuccess) return nullptr;
        if (cudaMalloc(&ws->psum, (size_t)cap * ITEMS * 8 * sizeof(float4)) != cudaSuccess) return nullptr;
        if (cudaMalloc(&ws->delta, (size_t)cap * 8 * sizeof(float4)) != cudaSuccess) return nullptr;
        ws->cap_units = cap;
    }
    return ws;
}

}  // namespace

// Returns -1 when the shape is not handled here (the caller goes on to mtm_tma.cu / mtm_dif.cu).
int mtm_launch_4s(int log2n, const MtmArgs& a, cudaStream_t stream) {
    static const int mode = getenv("SPYB_MTM_4S") ? atoi(getenv("SPYB_MTM_4S")) : 0;     // opt-in, see the header
    if (!mode || log2n != 12) return -1;
    if (a.out_kind != OUT_FOURIER_PLANAR || !a.keeptapers || a.n_tapers != 1) return -1;
    if (a.n_win != N || a.n_dft != N || a.n_frames != 1 || a.frame_start0 != 0 || a.n_samples < N) return -1;
    if (a.polyremoval > 0 || a.demean_taper || a.freq_idx != nullptr || a.n_freq_out != NF || a.chan_amax != nullptr) return -1;
    if (a.n_chan % 32 != 0 || !a.vec16 || a.n_trials < 1) return -1;
    if (a.so_trial % 4 || a.so_freq % 4 || reinterpret_cast<uintptr_t>(a.out) % 16 != 0) return -1;
    const long long n_units_ll = (long long)a.n_trials * (a.n_chan / 32);
    if (n_units_ll < 16 || n_units_ll > (1 << 24)) return -1;       // small launches: nothing to pipeline
    const int n_units = (int)n_units_ll;
    int dev = 0;
    SPYB_CUDA(cudaGetDevice(&dev));
    Workspace* ws = get_workspace(dev, stream, n_units);
    if (!ws) { cudaGetLastError(); return -1; }

    FsArgs f;
    f.x = a.x; f.trial_stride = a.trial_stride;
    f.n_chan = a.n_chan; f.groups = a.n_chan / 32; f.n_units = n_units;
    f.taper = a.tapers;
    f.half_scale = 0.5f * a.scale;
    f.detrend = a.polyremoval == 0 ? 1 : 0;
    f.out = static_cast<float*>(a.out);
    f.so_trial = a.so_trial; f.so_freq = a.so_freq;
    f.Y = ws->Y; f.ring_units = ws->ring_units; f.ub = ws->ub;
    f.n_batches = (n_units + ws->ub - 1) / ws->ub;
    f.cnt_m = ws->cnt; f.cnt_a = ws->cnt + ws->cap_units; f.cnt_b = ws->cnt + 2 * ws->cap_units;
    f.psum = ws->psum; f.delta = ws->delta; f.what2 = ws->what2;
    f.tw4096 = ws->tw4096;
    static const int dbg = getenv("SPYB_MTM_DBG") ? atoi(getenv("SPYB_MTM_DBG")) : 0;
    f.dbg = dbg;
    SPYB_CUDA(cudaMemsetAsync(ws->cnt, 0, (size_t)3 * ws->cap_units * sizeof(int), stream));
    if (f.detrend) {
        taper_phase_a_kernel<<<N / 256, 256, 0, stream>>>(a.tapers, ws->what2);
        SPYB_LAUNCH_CHECK("taper_phase_a_kernel");
        count_launch();
    }
    SPYB_CUDA(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    KERNEL<<<ws->grid, THREADS, SMEM_BYTES, stream>>>(f);
    SPYB_LAUNCH_CHECK("mtm_4s_kernel");
    count_launch();
    return 0;
}

}  // namespace spyb
