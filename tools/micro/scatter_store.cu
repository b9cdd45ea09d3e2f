// Micro-benchmark: how fast can one SM push scattered small pieces to global memory?
//   mode 0: STG.128, piece = P bytes (P/16 adjacent lanes), pieces S bytes apart
//   mode 1: cp.async.bulk.global.shared (one bulk copy per piece, issued by many threads)
//   mode 2: TMA tensor store (cp.async.bulk.tensor.2d), box = 32 rows x P bytes, rows S bytes apart
// 148 blocks x 512 threads; every block writes TILE = 128 KB per iteration, neighbouring blocks write neighbouring
// pieces (like the channel tiles of the tapered-FFT kernel).  Not part of the library.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(512, 1) k_stg(float4* out, int piece, long long stride_f4, int iters, int npieces_per_iter) {
    // piece bytes; lanes_per_piece = piece / 16
    const int lpp = piece / 16;
    const int tid = threadIdx.x;
    const int sub = tid % lpp, pidx = tid / lpp;          // piece handled by this thread in a round
    const int ppr = 512 / lpp;                            // pieces per round
    float4 v = make_float4(tid, 1.f, 2.f, 3.f);
    for (int it = 0; it < iters; ++it) {
        for (int p = pidx; p < npieces_per_iter; p += ppr) {
            // piece p of this block: row p (stride), column block blockIdx.x
            long long off = (long long)p * stride_f4 + (long long)(it % 8) * (gridDim.x * lpp) * 0 + (long long)blockIdx.x * lpp + sub;
            out[off + (long long)(it & 7) * 0] = v;
        }
        v.x += 1.f;
    }
}

__global__ void __launch_bounds__(512, 1) k_bulk(float4* out, int piece, long long stride_f4, int iters, int npieces_per_iter) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int tid = threadIdx.x;
    for (int i = tid; i < 32768 / 4; i += 512) reinterpret_cast<float*>(sm)[i] = (float)i;
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int lpp = piece / 16;
    for (int it = 0; it < iters; ++it) {
        for (int p = tid; p < npieces_per_iter; p += 512) {
            long long off = (long long)p * stride_f4 + (long long)blockIdx.x * lpp;
            const uint32_t src = smem_u32(sm + (p * piece) % 32768);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + off), "r"(src), "r"(piece) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(512, 1) k_tma(const __grid_constant__ CUtensorMap tmap, int piece, int iters, int npieces_per_iter, int rows_per_box) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int tid = threadIdx.x;
    for (int i = tid; i < 65536 / 4; i += 512) reinterpret_cast<float*>(sm)[i] = (float)i;
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int nbox = npieces_per_iter / rows_per_box;
    if (tid == 0) {
        for (int it = 0; it < iters; ++it) {
            for (int b = 0; b < nbox; ++b) {
                const uint32_t src = smem_u32(sm + (b * rows_per_box * piece) % 65536);
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                             ::"l"(reinterpret_cast<uint64_t>(&tmap)), "r"((int)(blockIdx.x * (piece / 4))), "r"(b * rows_per_box), "r"(src) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int iters = 44;
    const long long stride_bytes = 409600;         // 200 rows x 2 KB: distance between frequencies
    const size_t total = (size_t)4200 * stride_bytes;
    float4* out;
    CK(cudaMalloc(&out, total));
    CK(cudaMemset(out, 0, total));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeTiledFn encode = (EncodeTiledFn)fn;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (int piece : {32, 64, 128, 256, 512}) {
        const int npieces = 131072 / piece;        // 128 KB per block and iteration
        if (148 * piece > 8192) { /* pieces of all blocks must fit in one 'row' of stride bytes */ }
        for (int mode = 0; mode < 3; ++mode) {
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k_stg<<<148, 512>>>(out, piece, stride_bytes / 16, iters, npieces);
                else if (mode == 1) k_bulk<<<148, 512, 65536>>>(out, piece, stride_bytes / 16, iters, npieces);
                else {
                    CUtensorMap tmap;
                    const cuuint64_t gdim[2] = {(cuuint64_t)(stride_bytes / 4), 4200};
                    const cuuint64_t gstr[1] = {(cuuint64_t)stride_bytes};
                    const int rows = 128;
                    const cuuint32_t box[2] = {(cuuint32_t)(piece / 4), (cuuint32_t)rows};
                    const cuuint32_t es[2] = {1, 1};
                    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); break; }
                    k_tma<<<148, 512, 65536>>>(tmap, piece, iters, npieces, rows);
                }
                cudaEventRecord(e1);
                CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            const double bytes = 148.0 * iters * 131072;
            printf("piece %4d B mode %d (%s): %.3f ms  %.0f GB/s  %.2f us per 128 KB tile\n", piece, mode,
                   mode == 0 ? "STG.128" : mode == 1 ? "bulk copy" : "TMA store", best, bytes / best / 1e6, best * 1e3 / iters);
        }
    }
    return 0;
}
