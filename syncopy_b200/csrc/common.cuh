// Shared helpers for the syncopy_b200 CUDA engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace spyb {

// ---- thread-local error string handed out through spyb_last_error() -------------------
inline char* err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}
inline int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return 1;
}
#define SPYB_CUDA(call)                                                                        \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return spyb::fail("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,             \
                              cudaGetErrorString(_e));                                         \
    } while (0)
#define SPYB_LAUNCH_CHECK(name)                                                                \
    do {                                                                                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess)                                                                 \
            return spyb::fail("launch of %s failed: %s", name, cudaGetErrorString(_e));        \
    } while (0)

// ---- complex helpers (float2 = re, im) ------------------------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {   // a * conj(b)
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)
__device__ __forceinline__ float2 cmul_pi(float2 a) { return make_float2(-a.y, a.x); }   // a * (+i)

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

// upper-triangular 128 x 128 tiles of an nb x nb block matrix, row-major: t = 0 -> (0,0), 1 -> (0,1), ..., nb -> (1,1), ...
__host__ __device__ __forceinline__ int tri_tiles(int nb) { return nb * (nb + 1) / 2; }
__host__ __device__ __forceinline__ int tri_diag(int nb, int b) { return b * nb - b * (b - 1) / 2; }   // tile (b, b)
__host__ __device__ __forceinline__ void tri_decode(int nb, int t, int& ti, int& tj) {
    ti = 0;
    int row = nb;
    while (t >= row) { t -= row; ++ti; --row; }
    tj = ti + t;
}

// output conversions of syncopy/shared/const_def.py:25-40 (spectralConversions)
enum OutKind : int {
    OUT_POW = 0, OUT_ABS = 1, OUT_FOURIER = 2, OUT_REAL = 3, OUT_IMAG = 4,
    OUT_ANGLE = 5, OUT_ABSREAL = 6, OUT_ABSIMAG = 7,
    // engine-internal: complex result as two float32 planes, re at the element offset and im `n_chan`
    // floats later ([..][re|im][channel]); the operand layout of the tcgen05 cross-spectral kernel
    OUT_FOURIER_PLANAR = 8
};
__host__ __device__ __forceinline__ bool out_is_complex(int kind) { return kind == OUT_FOURIER; }

__device__ __forceinline__ float convert_real(float2 z, int kind) {
    switch (kind) {
        case OUT_POW:     return z.x * z.x + z.y * z.y;
        case OUT_ABS:     return hypotf(z.x, z.y);
        case OUT_REAL:    return z.x;
        case OUT_IMAG:    return z.y;
        case OUT_ANGLE:   return atan2f(z.y, z.x);
        case OUT_ABSREAL: return fabsf(z.x);
        default:          return fabsf(z.y);   // OUT_ABSIMAG
    }
}

}  // namespace spyb
