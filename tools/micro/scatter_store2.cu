// Micro-benchmark 2: the tapered-FFT kernel's result write pattern with its real footprint (839 MB written exactly
// once, DRAM write-back included).  Tile t = block + 148*it -> (trial = t / 32, ctile = t % 32); every tile writes
// 2049 bins x 2 planes x 32 bytes.  Layout A: [f][trial][plane][256 ch] (32-byte pieces); layout B:
// [f][trial][ctile][plane][8 ch] (64-byte pieces).  mode 0: STG.128 by 512 threads, mode 1: TMA tensor store from
// shared memory (one box = 128 bins x 2 planes x 8 channels).  Not part of the library.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int NF = 2048, R = 200, C = 256;
constexpr long long FSTRIDE = (long long)R * 2 * C;   // floats

__global__ void __launch_bounds__(512, 1) k_stg(float* out, int layoutB, int n_tiles, int spin) {
    const int tid = threadIdx.x, h = tid & 1, m = tid >> 1;      // m: 0..255
    float4 v = make_float4(tid, 1.f, 2.f, 3.f);
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int trial = t / 32, ct = t % 32;
        for (int s = 0; s < spin; ++s) v.x = v.x * 1.0001f + 0.5f;   // stand-in for compute between store bursts
        for (int g = 0; g < 8; ++g) {
            const int f = g * 256 + m;
            float* base = out + (long long)f * FSTRIDE + (long long)trial * 2 * C;
            if (!layoutB) {
                *reinterpret_cast<float4*>(base + ct * 8 + 4 * h) = v;
                *reinterpret_cast<float4*>(base + C + ct * 8 + 4 * h) = v;
            } else {
                *reinterpret_cast<float4*>(base + ct * 16 + 4 * h) = v;
                *reinterpret_cast<float4*>(base + ct * 16 + 8 + 4 * h) = v;
            }
        }
    }
    if (v.x == 1234.5f) out[0] = v.x;
}

__global__ void __launch_bounds__(512, 1) k_tma(const __grid_constant__ CUtensorMap tmap, int n_tiles) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int tid = threadIdx.x;
    for (int i = tid; i < 131072 / 4; i += 512) reinterpret_cast<float*>(sm)[i] = (float)i;
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) {
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int trial = t / 32, ct = t % 32;
            for (int k0 = 0; k0 < 16; ++k0) {
                const uint32_t src = smem_u32(sm + k0 * 8192);
                asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];"
                             ::"l"(reinterpret_cast<uint64_t>(&tmap)), "r"(ct * 8), "r"(0), "r"(0), "r"(k0), "r"(trial), "r"(src) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    const size_t total = (size_t)(NF + 1) * FSTRIDE * 4;
    float* out; CK(cudaMalloc(&out, total)); CK(cudaMemset(out, 0, total));
    float* flush; CK(cudaMalloc(&flush, 512u << 20));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeTiledFn encode = (EncodeTiledFn)fn;
    CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int n_tiles = R * 32;
    auto timeit = [&](const char* name, auto launch) {
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaMemset(flush, rep, 512u << 20));
            cudaEventRecord(e0); launch(); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        printf("%-44s %.3f ms  %.0f GB/s  %.2f us per tile and SM\n", name, best, n_tiles * 131072.0 / best / 1e6, best * 1e3 / (n_tiles / 148.0));
    };
    timeit("STG.128 layout A (32 B pieces)", [&] { k_stg<<<148, 512>>>(out, 0, n_tiles, 0); });
    timeit("STG.128 layout B (64 B pieces)", [&] { k_stg<<<148, 512>>>(out, 1, n_tiles, 0); });
    timeit("STG.128 layout A + ~7 us compute per tile", [&] { k_stg<<<148, 512>>>(out, 0, n_tiles, 3300); });
    timeit("STG.128 layout B + ~7 us compute per tile", [&] { k_stg<<<148, 512>>>(out, 1, n_tiles, 3300); });
    timeit("compute only (~7 us per tile)", [&] { k_stg<<<148, 512>>>(out, 0, 0 * n_tiles + 148 * 0, 3300); });
    for (int lay = 0; lay < 2; ++lay) {
        CUtensorMap tmap;
        // (c, m, plane, k0, trial): f = k0 + 16 m
        const cuuint64_t gdimA[5] = {C, 128, 2, 16, R};
        const cuuint64_t gstrA[4] = {(cuuint64_t)FSTRIDE * 16 * 4, C * 4, (cuuint64_t)FSTRIDE * 4, 2 * C * 4};
        const cuuint64_t gdimB[5] = {8, 128, 2, 16, (cuuint64_t)R * 32};          // B: c within block; 5th dim = trial*32 + ctile
        const cuuint64_t gstrB[4] = {(cuuint64_t)FSTRIDE * 16 * 4, 32, (cuuint64_t)FSTRIDE * 4, 64};
        const cuuint32_t box[5] = {8, 128, 2, 1, 1};
        const cuuint32_t es[5] = {1, 1, 1, 1, 1};
        CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, out, lay ? gdimB : gdimA, lay ? gstrB : gstrA, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
        if (lay == 0) timeit("TMA store layout A (32 B rows)", [&] { k_tma<<<148, 512, 131072>>>(tmap, n_tiles); });
        else {
            // layout B through the same kernel: coordinates (0, 0, 0, k0, trial*32+ct) -> reuse by passing ct*8 = 0: needs its own kernel; approximate with A-kernel semantics
            printf("(layout B TMA variant skipped)\n");
        }
    }
    // plain contiguous copy-out as the floor: every block writes its 128 KB tile contiguously
    return 0;
}
