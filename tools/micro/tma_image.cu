// Micro-test: which shared-memory image does a TMA tensor load produce when the box's inner dimension is 32 bytes
// and the swizzle mode is SWIZZLE_128B_ATOM_32B (vs the 128-byte inner box the cross-spectral kernel uses today)?
// Global tensor: rows x 128 floats, value = row * 1000 + col.  Not part of the library.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int ROWS = 16, COLS = 128;

__global__ void k(const __grid_constant__ CUtensorMap tmap, int mode, float* out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char* base = sm + ((1024u - (smem_u32(sm) & 1023u)) & 1023u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + 8192);
    float* tile = reinterpret_cast<float*>(base);
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) tile[i] = -1.f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(8192) : "memory");
        if (mode == 0)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(smem_u32(bar)), "r"(0), "r"(0), "r"(0) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(smem_u32(bar)), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
    }
    // wait
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(0) : "memory");
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = tile[i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn encode = (EncodeTiledFn)p;
    float* g; CK(cudaMalloc(&g, ROWS * COLS * 4));
    float h[ROWS * COLS];
    for (int r = 0; r < ROWS; ++r) for (int c = 0; c < COLS; ++c) h[r * COLS + c] = r * 1000.f + c;
    CK(cudaMemcpy(g, h, sizeof(h), cudaMemcpyHostToDevice));
    float* out; CK(cudaMalloc(&out, 8192));
    static float img[6][2048];
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
    const CUtensorMapSwizzle swz[6] = {CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_SWIZZLE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_SWIZZLE_64B};
    for (int mode = 0; mode < 6; ++mode) {
        if (only >= 0 && mode != only) continue;
        CUtensorMap tm;
        CUresult r;
        const cuuint32_t es[4] = {1, 1, 1, 1};
        if (mode == 0) {       // [chunk 4][row 16][32 floats]
            const cuuint64_t gd[3] = {32, ROWS, 4};
            const cuuint64_t gs[2] = {COLS * 4, 128};
            const cuuint32_t bx[3] = {32, ROWS, 4};
            r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, g, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz[mode],
                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {               // [chunk 4][row 16][4][8 floats]
            const cuuint64_t gd[4] = {8, 4, ROWS, 4};
            const cuuint64_t gs[3] = {32, COLS * 4, 128};
            const cuuint32_t bx[4] = {8, 4, ROWS, 4};
            r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz[mode],
                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (r != CUDA_SUCCESS) { printf("mode %d: encode failed %d\n", mode, (int)r); continue; }
        k<<<1, 128, 16384>>>(tm, mode == 0 ? 0 : 1, out);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(img[mode], out, 8192, cudaMemcpyDeviceToHost));
        printf("mode %d image, first 6 rows of 128 bytes (32-byte groups: first value of each):\n", mode);
        for (int line = 0; line < 8; ++line) {
            printf("  line %2d:", line);
            for (int ch = 0; ch < 4; ++ch) printf(" %8.0f", img[mode][line * 32 + ch * 8]);
            printf("\n");
        }
    }
    if (only >= 0) return 0;
    int same = memcmp(img[0], img[1], 8192) == 0;
    printf("image(32-byte inner, atom32 swizzle) == image(128-byte inner, atom32 swizzle): %s\n", same ? "YES" : "NO");
    // describe mode 1 as a permutation of mode 2 (unswizzled)
    int mism = 0;
    for (int i = 0; i < 2048 && mism < 8; ++i) if (img[0][i] != img[1][i]) { printf("  first diffs at float %d: %g vs %g\n", i, img[0][i], img[1][i]); ++mism; }
    return 0;
}
