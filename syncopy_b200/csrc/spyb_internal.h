// Internal interfaces between the translation units of libspyb200.so (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace spyb {

void count_launch();                 // bumps the counter behind spyb_launch_count()

// Generic "tapered real FFT of frames" job; mtmfft is the 1-frame case (see mtm.cu)
struct MtmFramesDesc {
    const float* x = nullptr;        // device, [trial][sample][channel]
    long long trial_stride = 0;
    int n_trials = 0, n_samples = 0, n_chan = 0;
    int n_win = 0, n_dft = 0;
    int frame_start0 = 0, hop = 1, n_frames = 1;
    const float* tapers = nullptr;   // device, [n_tapers][n_win]
    int n_tapers = 1;
    int polyremoval = -1, demean_taper = 0;
    float scale = 1.f;
    const int* freq_idx = nullptr;   // device or null
    int n_freq_out = 0;
    int out_kind = 0, keeptapers = 1;
    void* out = nullptr;             // device
    long long so_trial = 0, so_frame = 0, so_taper = 0, so_freq = 0;
    float* chan_amax = nullptr;
};
int mtm_frames(const MtmFramesDesc& d, cudaStream_t stream);
// fft_long.cu: transform lengths beyond the shared-memory kernels (global-memory Stockham passes, Bluestein)
bool mtm_needs_long(int n_dft);
int mtm_frames_long(const MtmFramesDesc& d, cudaStream_t stream);

// Cross-spectral contraction acc = beta*acc + alpha * sum_r X_r X_r^H per frequency (csd.cu)
struct CsdDesc {
    const void* spectra = nullptr;   // device complex64, element (f, r, c) at f*sx_f + r*sx_r + c
    long long sx_f = 0, sx_r = 0;
    int n_rows = 0, n_freq = 0, n_chan = 0;
    const int* idx_i = nullptr;      // optional sender / receiver channel subsets (device)
    const int* idx_j = nullptr;
    int n_i = 0, n_j = 0;
    float alpha = 1.f, beta = 0.f;
    void* acc = nullptr;             // device complex64 [n_freq][Ci][Cj]
};
int csd_accumulate_simt(const CsdDesc& d, cudaStream_t stream);

// Same contraction from planar spectra float32 [f][r][re|im][c] on tcgen05 tensor cores (csd_tc.cu)
struct CsdPlanarDesc {
    const float* planes = nullptr;   // device, element (f, r, plane, c) at f*sx_f + r*sx_r + plane*n_chan + c
    long long sx_f = 0, sx_r = 0;    // in floats
    int n_rows = 0, n_freq = 0, n_chan = 0;
    float alpha = 1.f, beta = 0.f;
    void* acc = nullptr;             // device complex64 [n_freq][C][C]
};
bool csd_tc_supported(int n_chan, long long sx_f, long long sx_r);
int csd_accumulate_tc(const CsdPlanarDesc& d, cudaStream_t stream);
// "tile slots": the upper 128x128 tiles of every frequency go, unmirrored, to the rank owning the frequency
// (peer-mapped base pointers; d.acc is ignored); csd_normalize_tiles sums the source ranks and normalises
// contraction + coherency in one kernel (one rank, all rows at once; d.acc, d.alpha, d.beta are ignored)
int csd_coherence_tc(const CsdPlanarDesc& d, int out_kind, void* out, cudaStream_t stream);
int csd_tile_count(int n_chan);
int csd_accumulate_tc_tiles(const CsdPlanarDesc& d, void* const* owner_base, const int* f_begin, int n_owners,
                            int src_rank, int skip_own, cudaStream_t stream);
int csd_coherence_tc_slots(const CsdPlanarDesc& d, const void* slots, int n_src, int skip_src, int out_kind, void* out,
                           cudaStream_t stream);
int csd_normalize_tiles(const void* slots, int n_src, int n_freq, int n_chan, float pre_scale, int out_kind,
                        void* out, cudaStream_t stream);
// Wavelet / superlet transforms as FFT convolutions (cwt.cu)
struct CwtDesc {
    int transposed = 0;              // 1: xspec [trial][chan][L/2+1], out [trial][scale][chan][n_time]
    const void* xspec = nullptr;     // device complex64 [trial][L/2+1][chan], spectra of the zero-padded trials
    int n_trials = 0, n_chan = 0, n_dft = 0;
    const void* kern = nullptr;      // device complex64 [scale][max_fac][L] = FFT_L(h) / L
    const float* expo = nullptr;     // device [scale][max_fac]
    const int* n_fac = nullptr;      // device [scale]
    int n_scales = 0, max_fac = 1;
    int n_time = 0, out_kind = 0;
    void* out = nullptr;             // device [trial][n_time][scale][chan]
};
int cwt_factors(const CwtDesc& d, cudaStream_t stream);
int cwt_factors_long(const CwtDesc& d, cudaStream_t stream);   // fft_long.cu: circular length > 16384
int transpose2d(const void* in, void* out, int batch, int rows, int cols, int elem_bytes, cudaStream_t stream);
int transpose_place(const void* in, void* out, int nb1, int nb2, int rows, int cols, int elem_bytes, long long stride_b1,
                    long long stride_b2, long long ld_out, int col_limit, cudaStream_t stream);
int detrend(const float* x, int n_trials, long long trial_stride, int n_samples, int n_chan, int polyremoval,
            float* out, long long out_trial_stride, cudaStream_t stream);
int gather_rows(const float* src, int n_trials, long long src_trial_stride, const int* idx, int n_idx,
                long long row_elems, float* dst, cudaStream_t stream);

int csd_normalize(const void* csd, long long n_mat, int n_chan, float pre_scale, int out_kind, void* out,
                  cudaStream_t stream);
int scale_inplace(float* x, long long n, float s, cudaStream_t stream);
int csd_mirror_upper(void* csd, int n_freq, int n_chan, cudaStream_t stream);
int sum_trials(const float* src, int n_trials, long long trial_stride, long long n_elems, float alpha, float beta,
               float* acc, cudaStream_t stream);

// Jackknife / PPC / cross-covariance helpers (stats.cu)
int axpby(const float* x, const float* y, float a, float b, float* out, long long n, cudaStream_t st);
int sqdev_accumulate(const float* avg, const float* x, float* var, long long n_elem, int is_complex, cudaStream_t st);
int unit_accumulate(const void* z, void* acc, long long n, int first, cudaStream_t st);
int ppc_finish(const void* acc, float* out, long long n, int n_trials, cudaStream_t st);
int xcov_kernel_spectra(const void* xspec, int n_chan, int L, int n, int shift, void* kern, cudaStream_t st);
int xcov_finish(const float* corr, const void* xspec, int n_chan, int n, int n_lags, int L, int norm, float* out,
                cudaStream_t st);

// Preprocessing (preproc.cu)
int sosfilt(const float* x, int n_trials, long long trial_stride, int n_samples, int n_chan, const double* sos_host,
            int n_sections, const double* zi_host, int edge, int twopass, double* scratch, float* out, cudaStream_t st);
int upfirdn(const float* x, int n_trials, long long trial_stride, int n_in, int n_chan, const double* h, int len_h,
            int up, int down, int m0, int n_out, float* out, cudaStream_t st);
int standardize(const float* x, int n_trials, long long trial_stride, int n_samples, int n_chan, float* out, cudaStream_t st);
int rectify(const float* x, float* out, long long n, cudaStream_t st);

// Granger path: regularisation, Wilson spectral factorisation, Geweke-Granger formula (wilson.cu); FP64
long long regularize_workspace_bytes(int n_freq, int n_chan);
int regularize_csd(const void* csd_c64, int n_freq, int n_chan, double cond_max, double eps_max, int n_steps,
                   void* out_c128, double* eps_host, double* cond0_host, void* work, long long work_bytes,
                   cudaStream_t stream);
long long wilson_workspace_bytes(int n_freq, int n_chan);
typedef int (*WilsonExchangeFn)(void* ctx, int what, void* buf, long long row_bytes, int n_rows);
int wilson_sf(const void* csd_c128, int n_freq, int n_chan, int n_iter, double rtol, void* H_out, double* Sigma_out,
              int* converged_host, double* err_host, int* iters_host, void* work, long long work_bytes,
              int f_lo, int f_hi, WilsonExchangeFn exchange, void* exchange_ctx, cudaStream_t stream);
int granger(const void* csd_c128, const void* H, const double* Sigma, int n_freq, int n_chan, float* out,
            cudaStream_t stream);

}  // namespace spyb
