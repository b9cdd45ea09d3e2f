// Host-side FFT plans: device-resident twiddle / chirp tables, computed in double precision
// and cached per length for the lifetime of the process (one CUDA context per process, see
// SURVEY.md 8b "Threading").
#pragma once
#include <cuda_runtime.h>

namespace spyb {

struct FftPlan {
    int n_dft = 0;        // logical DFT length (any n >= 1)
    int log2n = 0;        // block FFT length N = 2^log2n (== n_dft when direct)
    bool bluestein = false;
    const float2* tw = nullptr;      // pass twiddles for length N (Stockham passes, mtm.cu / cwt.cu)
    const float2* tw_dif = nullptr;  // pass twiddles of the in-place DIF passes (mtm_dif.cu), direct lengths only
    const float2* chirp = nullptr;   // bluestein: b[i] = exp(+i pi i^2 / n), i < n
    const float2* bhat = nullptr;    // bluestein: FFT_N(b wrapped) / N   (the 1/N of the inverse folded in)
};

// Returns nullptr and sets the error string if the length is unsupported.
const FftPlan* get_fft_plan(int n_dft);

// double-precision plan for the Wilson lag-domain transforms (power-of-two only)
struct FftPlanD {
    int n = 0, log2n = 0;
    const double2* tw = nullptr;     // W_n^k, k < n/2
};
const FftPlanD* get_fft_plan_d(int n);

void free_all_plans();

}  // namespace spyb
