// Kernel argument block shared by the two tapered-FFT kernels (mtm.cu: Stockham + Bluestein, mtm_dif.cu: in-place
// decimation-in-frequency for power-of-two lengths).
#pragma once
#include <cuda_runtime.h>

namespace spyb {

struct MtmArgs {
    const float* x;            // [trial][sample][channel]
    long long trial_stride;    // elements between trials
    int n_trials, n_samples, n_chan;
    int n_win;                 // detrend + taper window length (samples taken from the signal)
    int n_dft;                 // logical DFT length (>= n_win)
    int frame_start0, hop, n_frames;   // frame f starts at sample frame_start0 + f*hop (may be < 0: zeros)
    const float* tapers;       // [n_tapers][n_win]
    int n_tapers;
    int polyremoval;           // -1 none, 0 de-mean, 1 linear (over the n_win window, zeros included)
    int demean_taper;          // subtract the mean of the tapered window (mtmfft.py:114-116)
    float scale;               // spectrum scale (sqrt(2)/norm), the 1/2 of the pair split is folded in by the kernel
    const int* freq_idx;       // optional gather list (bins of the one-sided spectrum), may be null
    int n_freq_out;
    int out_kind, keeptapers;
    void* out;                 // float or float2 elements
    long long so_trial, so_frame, so_taper, so_freq;   // output strides in elements; channel stride is 1
    int vec_in, vec_out;       // alignment allows 8-byte input loads / paired output stores
    int vec16;                 // alignment allows 16-byte asynchronous row copies (n_chan % 4 == 0, 16-byte aligned base)
    const float2* tw_dif;      // pass twiddles of the in-place DIF passes
    int acc_smem;              // mtm_dif: the taper mean (keeptapers = 0) accumulates in shared memory, one store at the end
    float* chan_amax;          // optional [n_chan]: running max(|re|,|im|) of the scaled spectrum
    const float2* tw;
    const float2* chirp;
    const float2* bhat;
};

// mtm_dif.cu: returns -1 when the shape is not handled there (the caller then uses the Stockham kernel)
int mtm_launch_dif(int log2n, const MtmArgs& a, cudaStream_t stream);
// mtm_4s.cu: four-step kernel (32-channel units, 128-byte lines in and out) for N = 4096, one taper, planar output; -1 when not eligible
int mtm_launch_4s(int log2n, const MtmArgs& a, cudaStream_t stream);
// mtm_tma.cu: persistent TMA-pipelined kernel for N = 4096 and full 8-channel tiles; -1 when not eligible
int mtm_launch_tma(int log2n, const MtmArgs& a, cudaStream_t stream);
// mtm_r8.cu: 512-point windows (radix-8 passes, frame resident in registers across tapers); -1 when not eligible
int mtm_launch_r8(int log2n, const MtmArgs& a, cudaStream_t stream);

}  // namespace spyb
