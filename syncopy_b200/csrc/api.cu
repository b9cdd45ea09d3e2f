// extern "C" surface of libspyb200.so -- see include/spyb200.h for the contract.
#include "../../include/spyb200.h"
#include "common.cuh"
#include "plan.cuh"
#include "spyb_internal.h"

#include <atomic>

namespace spyb {
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace spyb

using namespace spyb;

extern "C" {

int spyb_version(void) { return SPYB_VERSION; }

const char* spyb_last_error(void) { return err_buf(); }

long long spyb_launch_count(void) { return g_launches.load(); }

int spyb_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail("no CUDA device available (%s)", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail("device %d out of range (have %d)", device, n);
    SPYB_CUDA(cudaSetDevice(device));
    SPYB_CUDA(cudaFree(0));
    cudaDeviceProp prop;
    SPYB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail("libspyb200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    return 0;
}

int spyb_max_fft_len(int pow2) { return pow2 ? 16384 : 8192; }

int spyb_mtmfft(const float* x, int n_trials, long long trial_stride, int n_samples, int n_chan,
                const float* tapers, int n_tapers, int nfft, float scale,
                int polyremoval, int demean_taper,
                const int* freq_idx, int n_freq_out, int out_kind, int keeptapers,
                void* out, long long so_trial, long long so_taper, long long so_freq,
                float* chan_amax, void* stream) {
    if (nfft < n_samples) return fail("nfft (%d) must be >= n_samples (%d)", nfft, n_samples);
    if (out_kind < 0 || out_kind > 8) return fail("bad out_kind %d", out_kind);
    MtmFramesDesc d;
    d.x = x; d.trial_stride = trial_stride;
    d.n_trials = n_trials; d.n_samples = n_samples; d.n_chan = n_chan;
    d.n_win = n_samples; d.n_dft = nfft;
    d.frame_start0 = 0; d.hop = 1; d.n_frames = 1;
    d.tapers = tapers; d.n_tapers = n_tapers;
    d.polyremoval = polyremoval; d.demean_taper = demean_taper; d.scale = scale;
    d.freq_idx = freq_idx; d.n_freq_out = n_freq_out;
    d.out_kind = out_kind; d.keeptapers = keeptapers;
    d.out = out; d.so_trial = so_trial; d.so_frame = 0; d.so_taper = so_taper; d.so_freq = so_freq;
    d.chan_amax = chan_amax;
    return mtm_frames(d, static_cast<cudaStream_t>(stream));
}

int spyb_mtmconvol(const float* x, int n_trials, long long trial_stride, int n_samples, int n_chan,
                   const float* tapers, int n_tapers, int nperseg, int hop, int frame_start0, int n_frames,
                   float scale, int polyremoval,
                   const int* freq_idx, int n_freq_out, int out_kind, int keeptapers,
                   void* out, long long so_trial, long long so_frame, long long so_taper, long long so_freq,
                   void* stream) {
    if (nperseg < 1 || hop < 1) return fail("nperseg (%d) and hop (%d) must be positive", nperseg, hop);
    if (out_kind < 0 || out_kind > 7) return fail("bad out_kind %d", out_kind);
    MtmFramesDesc d;
    d.x = x; d.trial_stride = trial_stride;
    d.n_trials = n_trials; d.n_samples = n_samples; d.n_chan = n_chan;
    d.n_win = nperseg; d.n_dft = nperseg;
    d.frame_start0 = frame_start0; d.hop = hop; d.n_frames = n_frames;
    d.tapers = tapers; d.n_tapers = n_tapers;
    d.polyremoval = polyremoval; d.demean_taper = 0; d.scale = scale;
    d.freq_idx = freq_idx; d.n_freq_out = n_freq_out;
    d.out_kind = out_kind; d.keeptapers = keeptapers;
    d.out = out; d.so_trial = so_trial; d.so_frame = so_frame; d.so_taper = so_taper; d.so_freq = so_freq;
    return mtm_frames(d, static_cast<cudaStream_t>(stream));
}

int spyb_csd_accumulate(const void* spectra, long long sx_f, long long sx_r, int n_rows, int n_freq, int n_chan,
                        const int* idx_i, int n_i, const int* idx_j, int n_j,
                        float alpha, float beta, void* acc, int impl, void* stream) {
    if ((idx_i == nullptr) != (idx_j == nullptr))
        return fail("idx_i and idx_j must both be given or both be NULL");
    CsdDesc d;
    d.spectra = spectra; d.sx_f = sx_f; d.sx_r = sx_r;
    d.n_rows = n_rows; d.n_freq = n_freq; d.n_chan = n_chan;
    d.idx_i = idx_i; d.idx_j = idx_j; d.n_i = n_i; d.n_j = n_j;
    d.alpha = alpha; d.beta = beta; d.acc = acc;
    if (impl == 2) return fail("the tcgen05 kernel takes planar spectra: call spyb_csd_accumulate_planar");
    return csd_accumulate_simt(d, static_cast<cudaStream_t>(stream));
}

int spyb_csd_planar_supported(int n_chan, long long sx_f, long long sx_r) {
    return csd_tc_supported(n_chan, sx_f, sx_r) ? 1 : 0;
}

int spyb_csd_accumulate_planar(const float* planes, long long sx_f, long long sx_r, int n_rows, int n_freq,
                               int n_chan, float alpha, float beta, void* acc, void* stream) {
    CsdPlanarDesc d;
    d.planes = planes; d.sx_f = sx_f; d.sx_r = sx_r;
    d.n_rows = n_rows; d.n_freq = n_freq; d.n_chan = n_chan;
    d.alpha = alpha; d.beta = beta; d.acc = acc;
    return csd_accumulate_tc(d, static_cast<cudaStream_t>(stream));
}

int spyb_csd_tile_count(int n_chan) { return csd_tile_count(n_chan); }

int spyb_csd_coherence_planar(const float* planes, long long sx_f, long long sx_r, int n_rows, int n_freq,
                              int n_chan, int out_kind, void* out, void* stream) {
    if (out_kind < 0 || out_kind > 7) return fail("bad out_kind %d", out_kind);
    CsdPlanarDesc d;
    d.planes = planes; d.sx_f = sx_f; d.sx_r = sx_r;
    d.n_rows = n_rows; d.n_freq = n_freq; d.n_chan = n_chan;
    d.alpha = 1.f; d.beta = 0.f; d.acc = nullptr;
    return csd_coherence_tc(d, out_kind, out, static_cast<cudaStream_t>(stream));
}

int spyb_csd_accumulate_tiles(const float* planes, long long sx_f, long long sx_r, int n_rows, int n_freq,
                              int n_chan, float alpha, float beta, void* const* owner_base_host,
                              const int* f_begin_host, int n_owners, int src_rank, void* stream) {
    if (!owner_base_host || !f_begin_host) return fail("owner_base_host / f_begin_host must not be NULL");
    CsdPlanarDesc d;
    d.planes = planes; d.sx_f = sx_f; d.sx_r = sx_r;
    d.n_rows = n_rows; d.n_freq = n_freq; d.n_chan = n_chan;
    d.alpha = alpha; d.beta = beta; d.acc = nullptr;
    return csd_accumulate_tc_tiles(d, owner_base_host, f_begin_host, n_owners, src_rank, 0,
                                   static_cast<cudaStream_t>(stream));
}

int spyb_csd_accumulate_tiles_others(const float* planes, long long sx_f, long long sx_r, int n_rows, int n_freq,
                                     int n_chan, float alpha, float beta, void* const* owner_base_host,
                                     const int* f_begin_host, int n_owners, int src_rank, void* stream) {
    if (!owner_base_host || !f_begin_host) return fail("owner_base_host / f_begin_host must not be NULL");
    CsdPlanarDesc d;
    d.planes = planes; d.sx_f = sx_f; d.sx_r = sx_r;
    d.n_rows = n_rows; d.n_freq = n_freq; d.n_chan = n_chan;
    d.alpha = alpha; d.beta = beta; d.acc = nullptr;
    return csd_accumulate_tc_tiles(d, owner_base_host, f_begin_host, n_owners, src_rank, 1,
                                   static_cast<cudaStream_t>(stream));
}

int spyb_csd_coherence_planar_slots(const float* planes, long long sx_f, long long sx_r, int n_rows, int n_freq,
                                    int n_chan, const void* slots, int n_src, int skip_src, int out_kind, void* out,
                                    void* stream) {
    if (out_kind < 0 || out_kind > 7) return fail("bad out_kind %d", out_kind);
    CsdPlanarDesc d;
    d.planes = planes; d.sx_f = sx_f; d.sx_r = sx_r;
    d.n_rows = n_rows; d.n_freq = n_freq; d.n_chan = n_chan;
    d.alpha = 1.f; d.beta = 0.f; d.acc = nullptr;
    return csd_coherence_tc_slots(d, slots, n_src, skip_src, out_kind, out, static_cast<cudaStream_t>(stream));
}

int spyb_csd_normalize_tiles(const void* slots, int n_src, int n_freq_local, int n_chan, float pre_scale,
                             int out_kind, void* out, void* stream) {
    if (out_kind < 0 || out_kind > 7) return fail("bad out_kind %d", out_kind);
    return csd_normalize_tiles(slots, n_src, n_freq_local, n_chan, pre_scale, out_kind, out,
                               static_cast<cudaStream_t>(stream));
}

int spyb_peer_alloc(long long bytes, void** ptr_out, unsigned char* handle64_out) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!ptr_out || !handle64_out || bytes <= 0) return fail("spyb_peer_alloc: bad arguments");
    void* p = nullptr;
    SPYB_CUDA(cudaMalloc(&p, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return fail("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); }
    memcpy(handle64_out, &h, 64);
    *ptr_out = p;
    return 0;
}

int spyb_peer_open(const unsigned char* handle64, void** ptr_out) {
    if (!ptr_out || !handle64) return fail("spyb_peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    SPYB_CUDA(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int spyb_peer_memset(void* ptr, int value, long long bytes, void* stream) {
    if (!ptr || bytes < 0) return fail("spyb_peer_memset: bad arguments");
    if (bytes == 0) return 0;
    SPYB_CUDA(cudaMemsetAsync(ptr, value, (size_t)bytes, static_cast<cudaStream_t>(stream)));
    return 0;
}

int spyb_peer_close(void* mapped_ptr) {
    SPYB_CUDA(cudaIpcCloseMemHandle(mapped_ptr));
    return 0;
}

int spyb_peer_free(void* ptr) {
    SPYB_CUDA(cudaFree(ptr));
    return 0;
}

int spyb_csd_normalize(const void* csd, long long n_mat, int n_chan, float pre_scale, int out_kind,
                       void* out, void* stream) {
    if (out_kind < 0 || out_kind > 7) return fail("bad out_kind %d", out_kind);
    return csd_normalize(csd, n_mat, n_chan, pre_scale, out_kind, out, static_cast<cudaStream_t>(stream));
}

int spyb_detrend(const float* x, int n_trials, long long trial_stride, int n_samples, int n_chan, int polyremoval,
                 float* out, long long out_trial_stride, void* stream) {
    return detrend(x, n_trials, trial_stride, n_samples, n_chan, polyremoval, out, out_trial_stride,
                   static_cast<cudaStream_t>(stream));
}

int spyb_transpose(const void* in, void* out, int batch, int rows, int cols, int elem_bytes, void* stream) {
    return transpose2d(in, out, batch, rows, cols, elem_bytes, static_cast<cudaStream_t>(stream));
}

int spyb_transpose_place(const void* in, void* out, int nb1, int nb2, int rows, int cols, int elem_bytes,
                         long long stride_b1, long long stride_b2, long long ld_out, int col_limit, void* stream) {
    return transpose_place(in, out, nb1, nb2, rows, cols, elem_bytes, stride_b1, stride_b2, ld_out, col_limit,
                           static_cast<cudaStream_t>(stream));
}

int spyb_cwt(const void* xspec, int n_trials, int n_chan, int n_dft, const void* kern, const float* expo,
             const int* n_fac, int n_scales, int max_fac, int n_time, int out_kind, int transposed, void* out,
             void* stream) {
    if (out_kind < 0 || out_kind > 7) return fail("bad out_kind %d", out_kind);
    CwtDesc d;
    d.transposed = transposed ? 1 : 0;
    d.xspec = xspec; d.n_trials = n_trials; d.n_chan = n_chan; d.n_dft = n_dft;
    d.kern = kern; d.expo = expo; d.n_fac = n_fac; d.n_scales = n_scales; d.max_fac = max_fac;
    d.n_time = n_time; d.out_kind = out_kind; d.out = out;
    return cwt_factors(d, static_cast<cudaStream_t>(stream));
}

int spyb_gather_rows(const float* src, int n_trials, long long src_trial_stride, const int* idx, int n_idx,
                     long long row_elems, float* dst, void* stream) {
    return gather_rows(src, n_trials, src_trial_stride, idx, n_idx, row_elems, dst, static_cast<cudaStream_t>(stream));
}

int spyb_csd_mirror_upper(void* csd, int n_freq, int n_chan, void* stream) {
    return csd_mirror_upper(csd, n_freq, n_chan, static_cast<cudaStream_t>(stream));
}

int spyb_scale(float* x, long long n, float s, void* stream) {
    return scale_inplace(x, n, s, static_cast<cudaStream_t>(stream));
}

int spyb_sum_trials(const float* src, int n_trials, long long trial_stride, long long n_elems, float alpha, float beta,
                    float* acc, void* stream) {
    return sum_trials(src, n_trials, trial_stride, n_elems, alpha, beta, acc, static_cast<cudaStream_t>(stream));
}

int spyb_axpby(const float* x, const float* y, float a, float b, float* out, long long n, void* stream) {
    return axpby(x, y, a, b, out, n, static_cast<cudaStream_t>(stream));
}

int spyb_sqdev_accumulate(const float* avg, const float* x, float* var, long long n_elem, int is_complex, void* stream) {
    return sqdev_accumulate(avg, x, var, n_elem, is_complex, static_cast<cudaStream_t>(stream));
}

int spyb_unit_accumulate(const void* z, void* acc, long long n, int first, void* stream) {
    return unit_accumulate(z, acc, n, first, static_cast<cudaStream_t>(stream));
}

int spyb_ppc_finish(const void* acc, float* out, long long n, int n_trials, void* stream) {
    return ppc_finish(acc, out, n, n_trials, static_cast<cudaStream_t>(stream));
}

int spyb_xcov_kernel_spectra(const void* xspec, int n_chan, int n_dft, int n_samples, int shift, void* kern, void* stream) {
    return xcov_kernel_spectra(xspec, n_chan, n_dft, n_samples, shift, kern, static_cast<cudaStream_t>(stream));
}

int spyb_xcov_finish(const float* corr, const void* xspec, int n_chan, int n_samples, int n_lags, int n_dft, int norm,
                     float* out, void* stream) {
    return xcov_finish(corr, xspec, n_chan, n_samples, n_lags, n_dft, norm, out, static_cast<cudaStream_t>(stream));
}

int spyb_sosfilt(const float* x, int n_trials, long long trial_stride, int n_samples, int n_chan, const double* sos_host,
                 int n_sections, const double* zi_host, int edge, int twopass, double* scratch, float* out, void* stream) {
    if (!sos_host) return fail("sos_host must not be NULL");
    return sosfilt(x, n_trials, trial_stride, n_samples, n_chan, sos_host, n_sections, zi_host, edge, twopass, scratch, out,
                   static_cast<cudaStream_t>(stream));
}

int spyb_upfirdn(const float* x, int n_trials, long long trial_stride, int n_in, int n_chan, const double* h, int len_h,
                 int up, int down, int first_row, int n_out, float* out, void* stream) {
    return upfirdn(x, n_trials, trial_stride, n_in, n_chan, h, len_h, up, down, first_row, n_out, out,
                   static_cast<cudaStream_t>(stream));
}

int spyb_standardize(const float* x, int n_trials, long long trial_stride, int n_samples, int n_chan, float* out, void* stream) {
    return standardize(x, n_trials, trial_stride, n_samples, n_chan, out, static_cast<cudaStream_t>(stream));
}

int spyb_rectify(const float* x, float* out, long long n, void* stream) {
    return rectify(x, out, n, static_cast<cudaStream_t>(stream));
}

long long spyb_regularize_workspace_bytes(int n_freq, int n_chan) {
    return regularize_workspace_bytes(n_freq, n_chan);
}

int spyb_regularize_csd(const void* csd, int n_freq, int n_chan, double cond_max, double eps_max, int n_steps,
                        void* out, double* eps_host, double* cond0_host, void* work, long long work_bytes,
                        void* stream) {
    if (!eps_host || !cond0_host) return fail("eps_host / cond0_host must not be NULL");
    return regularize_csd(csd, n_freq, n_chan, cond_max, eps_max, n_steps, out, eps_host, cond0_host, work,
                          work_bytes, static_cast<cudaStream_t>(stream));
}

long long spyb_wilson_workspace_bytes(int n_freq, int n_chan) { return wilson_workspace_bytes(n_freq, n_chan); }

int spyb_wilson(const void* csd, int n_freq, int n_chan, int n_iter, double rtol, void* H, double* Sigma,
                int* converged_host, double* err_host, int* iters_host, void* work, long long work_bytes,
                void* stream) {
    if (!converged_host || !err_host) return fail("converged_host / err_host must not be NULL");
    return wilson_sf(csd, n_freq, n_chan, n_iter, rtol, H, Sigma, converged_host, err_host, iters_host, work,
                     work_bytes, 0, n_freq, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

int spyb_wilson_sharded(const void* csd, int n_freq, int n_chan, int n_iter, double rtol, void* H, double* Sigma,
                        int* converged_host, double* err_host, int* iters_host, void* work, long long work_bytes,
                        int f_lo, int f_hi, spyb_exchange_fn exchange, void* exchange_ctx, void* stream) {
    if (!converged_host || !err_host) return fail("converged_host / err_host must not be NULL");
    if (!exchange) return fail("spyb_wilson_sharded needs an exchange callback (use spyb_wilson on one rank)");
    return wilson_sf(csd, n_freq, n_chan, n_iter, rtol, H, Sigma, converged_host, err_host, iters_host, work,
                     work_bytes, f_lo, f_hi, exchange, exchange_ctx, static_cast<cudaStream_t>(stream));
}

int spyb_granger(const void* csd, const void* H, const double* Sigma, int n_freq, int n_chan, float* out,
                 void* stream) {
    return granger(csd, H, Sigma, n_freq, n_chan, out, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
