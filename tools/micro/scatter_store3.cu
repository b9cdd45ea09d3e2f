// Micro-benchmark 3: why the tapered-FFT kernel's result stores do not overlap its passes, and what wider pieces
// would buy.  Same footprint as scatter_store2 (839 MB written once, [f][trial][plane][256 ch]); per tile a
// "compute" phase (none / register-only ALU / shared-memory passes like the FFT: 16 x LDS.128 + 16 x STS.128 +
// barrier) followed by the result stores in one of four patterns:
//   0  32-byte pieces, STG.128 by lane pairs (what the kernel does)
//   1  32-byte pieces, one STG.256 per lane (st.global.v8.f32)
//   2  128-byte pieces (a tile = 32 channels x a quarter of the bins: what a wider decomposition would write)
//   3  contiguous 128 KB per tile (floor)
// Not part of the library.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int NF = 2048, R = 200, C = 256;
constexpr long long FSTRIDE = (long long)R * 2 * C;   // floats

__device__ __forceinline__ void st_v8(float* p, float4 a, float4 b) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w),
                 "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}

__global__ void __launch_bounds__(512, 1) k_mix(float* out, int n_tiles, int pattern, int compute, int iters) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int tid = threadIdx.x, h = tid & 1, m = tid >> 1;
    float4 v = make_float4(tid, 1.f, 2.f, 3.f);
    float4* s4 = reinterpret_cast<float4*>(sm);
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        if (compute == 1) {
            for (int s = 0; s < iters; ++s) v.x = v.x * 1.0001f + 0.5f;
        } else if (compute == 2) {
            for (int pass = 0; pass < iters; ++pass) {
                float4 x[16];
#pragma unroll
                for (int r = 0; r < 16; ++r) x[r] = s4[(tid + r * 512) & 8191];
#pragma unroll
                for (int r = 0; r < 16; ++r) { x[r].x += v.x; x[r].y = x[r].y * 1.001f + x[(r + 1) & 15].z; }
#pragma unroll
                for (int r = 0; r < 16; ++r) s4[(tid + r * 512) & 8191] = x[r];
                __syncthreads();
                v.y += x[3].y;
            }
        }
        if (pattern == 0) {
            const int trial = t / 32, ct = t % 32;
            for (int g = 0; g < 8; ++g) {
                const int f = g * 256 + m;
                float* base = out + (long long)f * FSTRIDE + (long long)trial * 2 * C;
                *reinterpret_cast<float4*>(base + ct * 8 + 4 * h) = v;
                *reinterpret_cast<float4*>(base + C + ct * 8 + 4 * h) = v;
            }
        } else if (pattern == 1) {
            const int trial = t / 32, ct = t % 32;
            for (int g = 0; g < 4; ++g) {
                const int f = g * 512 + tid;
                float* base = out + (long long)f * FSTRIDE + (long long)trial * 2 * C;
                st_v8(base + ct * 8, v, v);
                st_v8(base + C + ct * 8, v, v);
            }
        } else if (pattern == 2) {
            const int trial = t / 32, ct32 = (t % 32) >> 2, fq = t & 3;
            const int chunk = tid & 7, b = tid >> 3;             // 64 bins per step
            for (int g = 0; g < 8; ++g) {
                const int f = fq * 512 + g * 64 + b;
                float* base = out + (long long)f * FSTRIDE + (long long)trial * 2 * C + ct32 * 32 + chunk * 4;
                *reinterpret_cast<float4*>(base) = v;
                *reinterpret_cast<float4*>(base + C) = v;
            }
        } else if (pattern == 3) {
            float4* base = reinterpret_cast<float4*>(out) + (long long)t * 8192;
            for (int g = 0; g < 16; ++g) base[g * 512 + tid] = v;
        }
    }
    if (v.x == 1234.5f) out[0] = v.x + v.y;
}

int main() {
    const size_t total = (size_t)(NF + 1) * FSTRIDE * 4;
    float* out; CK(cudaMalloc(&out, total)); CK(cudaMemset(out, 0, total));
    float* flush; CK(cudaMalloc(&flush, 512u << 20));
    CK(cudaFuncSetAttribute(k_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int n_tiles = R * 32;
    auto timeit = [&](int pattern, int compute, int iters, int tiles) {
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaMemset(flush, rep, 512u << 20));
            cudaEventRecord(e0); k_mix<<<148, 512, 131072>>>(out, tiles, pattern, compute, iters); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        return best;
    };
    const char* pn[5] = {"32 B pieces, STG.128 pairs", "32 B pieces, STG.256", "128 B pieces", "contiguous", "no stores"};
    // compute-only references (stores disabled by a pattern that writes nothing: use 0 tiles worth -> scale)
    for (int compute = 0; compute < 3; ++compute) {
        const int iters = compute == 1 ? 3300 : compute == 2 ? 6 : 0;
        const char* cn = compute == 0 ? "stores only" : compute == 1 ? "+ ALU-only compute" : "+ shared-memory passes";
        for (int p = 0; p < 5; ++p) {
            const float ms = timeit(p, compute, iters, n_tiles);
            printf("%-30s %-26s %.3f ms  %.2f us per tile and SM\n", pn[p], cn, ms, ms * 1e3 / (n_tiles / 148.0));
        }
    }
    return 0;
}
