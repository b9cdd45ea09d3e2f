/*
 * libspyb200 -- C ABI of the B200-native per-trial spectral / cross-spectral engine that
 * stands in for the NumPy/SciPy bodies behind Syncopy's `computeFunction`s.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in `_host`;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; 0 = default);
 *   - return value 0 = ok, non-zero = error, text via spyb_last_error() (thread local);
 *   - trials are C-contiguous [trial][sample][channel] float32 (AnalogData default dimord,
 *     reference syncopy/datatype/continuous_data.py:405); complex64 is interleaved (re, im);
 *   - `out_kind` follows syncopy/shared/const_def.py:25-40:
 *       0 pow, 1 abs, 2 fourier/complex, 3 real, 4 imag, 5 angle, 6 absreal, 7 absimag;
 *     spyb_mtmfft additionally accepts 8 = complex result as two float32 planes (re at the element
 *     offset, im n_chan floats later; output strides then count floats) -- the operand layout of
 *     spyb_csd_accumulate_planar.
 *
 * The reference has no FFI (it is pure Python); the interface each entry point replaces is
 * the NumPy-level function cited next to it.  INTEGRATION.md shows the ctypes binding and the
 * `computeFunction` shim a Syncopy maintainer would add.
 */
#ifndef SPYB200_H
#define SPYB200_H

#ifdef __cplusplus
extern "C" {
#endif

#define SPYB_VERSION 100

/* library / device management ---------------------------------------------------------- */
int spyb_version(void);
int spyb_init(int device);                 /* cudaSetDevice + warm-up of the context       */
const char* spyb_last_error(void);
long long spyb_launch_count(void);         /* number of kernels this library has launched  */
int spyb_max_fft_len(int pow2);            /* longest transform of the shared-memory kernels (pow2 / arbitrary);
                                              longer ones (up to 2^24) run as global-memory passes           */

/*
 * (Multi-)tapered FFT of whole trials.
 * Replaces: syncopy/specest/mtmfft.py:16-129 (`mtmfft`), plus the arithmetic of
 *           syncopy/specest/compRoutines.py:169-189 (`mtmfft_cF`: detrend, frequency gather,
 *           spectralConversions, taper mean).
 *   x            [n_trials][n_samples][n_chan] float32, `trial_stride` elements between trials
 *   tapers       [n_tapers][n_samples] float32, already normalised (_norm_spec.py:27-46)
 *   nfft         padded length (`nSamples` of the reference), any length >= n_samples
 *   scale        sqrt(2) / norm of `_norm_spec` (_norm_spec.py:22, mtmfft.py:119-127)
 *   polyremoval  -1 none, 0 de-mean, 1 linear detrend (scipy.signal.detrend over the trial)
 *   demean_taper subtract the mean of each tapered series (mtmfft.py:114-116)
 *   freq_idx     optional gather list into the nfft/2+1 bins (best_match, compRoutines.py:158)
 *   keeptapers   0: average the converted output over tapers (compRoutines.py:188-189)
 *   out          element (trial, taper, fi, chan) at trial*so_trial + taper*so_taper + fi*so_freq + chan;
 *                float32 or complex64 depending on out_kind
 *   chan_amax    optional [n_chan] float32, atomically max-updated with max(|re|,|im|)
 */
int spyb_mtmfft(const float* x, int n_trials, long long trial_stride, int n_samples, int n_chan,
                const float* tapers, int n_tapers, int nfft, float scale,
                int polyremoval, int demean_taper,
                const int* freq_idx, int n_freq_out, int out_kind, int keeptapers,
                void* out, long long so_trial, long long so_taper, long long so_freq,
                float* chan_amax, void* stream);

/*
 * (Multi-)tapered short-time FFT on sliding frames, no frames are materialised.
 * Replaces: syncopy/specest/stft.py:16-159 and syncopy/specest/mtmconvol.py:17-152.
 *   frame f covers samples [frame_start0 + f*hop, +nperseg) of the trial; samples outside
 *   [0, n_samples) read as zero (boundary='zeros' -> frame_start0 = -nperseg/2; the end padding
 *   of stft.py:112-117 is implied).  Per-segment detrending includes those zeros
 *   (stft.py:131-132).  tapers [n_tapers][nperseg]; scale = sqrt(2)/nperseg (stft.py:154).
 *   out element (trial, frame, taper, fi, chan) at
 *       trial*so_trial + frame*so_frame + taper*so_taper + fi*so_freq + chan.
 */
int spyb_mtmconvol(const float* x, int n_trials, long long trial_stride, int n_samples, int n_chan,
                   const float* tapers, int n_tapers, int nperseg, int hop, int frame_start0, int n_frames,
                   float scale, int polyremoval,
                   const int* freq_idx, int n_freq_out, int out_kind, int keeptapers,
                   void* out, long long so_trial, long long so_frame, long long so_taper, long long so_freq,
                   void* stream);

/*
 * Cross-spectral contraction:  acc[f][i][j] = beta*acc[f][i][j]
 *                                 + alpha * sum_{r < n_rows} X[f][r][si(i)] * conj(X[f][r][sj(j)])
 * Replaces: syncopy/connectivity/csd.py:98-102 (outer product + taper mean),
 *           syncopy/connectivity/ST_compRoutines.py:84-116 (`spectral_dyadic_product_cF`) and, when the
 *           rows of many trials are passed at once, the trial sum of
 *           syncopy/shared/computational_routine.py:1022-1025.
 *   spectra  complex64, element (f, r, c) at f*sx_f + r*sx_r + c   (r = trial*n_tapers + taper)
 *   idx_i/j  optional channel subsets (send_idx / rec_idx); NULL = all channels (Hermitian result)
 *   acc      complex64 [n_freq][Ci][Cj]
 *   impl     0 auto, 1 CUDA-core FP32 kernel, 2 tcgen05 tensor-core kernel (errors if not applicable)
 */
int spyb_csd_accumulate(const void* spectra, long long sx_f, long long sx_r, int n_rows, int n_freq, int n_chan,
                        const int* idx_i, int n_i, const int* idx_j, int n_j,
                        float alpha, float beta, void* acc, int impl, void* stream);

/*
 * The same contraction on the tcgen05 tensor cores (3xTF32 split, FP32 accumulation in TMEM), for the
 * whole-dataset path where all (trial, taper) rows of a frequency are reduced in one launch
 * (csd.py:98-102 + computational_routine.py:1022-1025).
 *   planes   float32 "planar" spectra as written by spyb_mtmfft with out_kind = 8:
 *            element (f, r, plane, c) at f*sx_f + r*sx_r + plane*n_chan + c, plane 0 = real, 1 = imaginary
 *   acc      complex64 [n_freq][n_chan][n_chan], acc = beta*acc + alpha * sum_r X_r X_r^H
 * spyb_csd_planar_supported() tells whether a shape is eligible (64 <= n_chan <= 512 in multiples of 32, strides % 4 == 0);
 * ineligible shapes go through spyb_csd_accumulate.
 */
int spyb_csd_planar_supported(int n_chan, long long sx_f, long long sx_r);
int spyb_csd_accumulate_planar(const float* planes, long long sx_f, long long sx_r, int n_rows, int n_freq,
                               int n_chan, float alpha, float beta, void* acc, void* stream);

/*
 * Sharded variant of the two calls above for the trial-averaged path (keeptrials=False,
 * computational_routine.py:1022-1032 + csd.py:118-172) on one or several GPUs.  Frequencies are split into
 * contiguous slabs, slab o = [f_begin_host[o], f_begin_host[o+1]) is owned by rank o.  spyb_csd_accumulate_tiles
 * writes every upper-triangular 128x128 tile of frequency f (tile order (0,0), (0,1), (1,1); one tile for 128
 * channels), scaled by alpha and unmirrored, into the owner's slot buffer
 *     owner_base_host[o] + ((src_rank * nF_o + (f - f_begin[o])) * n_tiles + tile) * 128*128   (complex64)
 * with plain stores -- owner_base_host[o] may be a peer-mapped pointer (CUDA IPC / NVLink P2P), so the transfer
 * of a finished tile overlaps the tensor-core work on the next one and no collective moves the data.  After a
 * barrier between the ranks, spyb_csd_normalize_tiles sums the n_src source slots of the local slab
 * (slots [n_src][nF_local][n_tiles][128][128]), multiplies by pre_scale (1/nTrials), normalises and writes
 * out [nF_local][n_chan][n_chan] (float32, or complex64 for out_kind 2) including the mirrored half.
 * With one rank (n_owners = n_src = 1) the pair replaces spyb_csd_accumulate_planar + spyb_csd_normalize without
 * ever materialising the mirrored CSD.  Eligibility as spyb_csd_planar_supported.
 */
int spyb_csd_tile_count(int n_chan);
/*
 * One rank, all (trial, taper) rows in one launch: cross-spectral sum (csd.py:98-102, computational_routine.py:1022-1032),
 * coherency normalisation and output conversion (csd.py:118-172) in ONE kernel -- the epilogue of the tcgen05
 * contraction divides by sqrt(C_ii C_jj) and writes out [n_freq][n_chan][n_chan] (float32, or complex64 for
 * out_kind 2) including the mirrored half; the cross-spectral matrix never reaches memory (coherency does not
 * depend on the 1/nTapers, 1/nTrials factors).  Same eligibility as spyb_csd_planar_supported.
 */
int spyb_csd_coherence_planar(const float* planes, long long sx_f, long long sx_r, int n_rows, int n_freq,
                              int n_chan, int out_kind, void* out, void* stream);
int spyb_csd_accumulate_tiles(const float* planes, long long sx_f, long long sx_r, int n_rows, int n_freq,
                              int n_chan, float alpha, float beta, void* const* owner_base_host,
                              const int* f_begin_host, int n_owners, int src_rank, void* stream);
int spyb_csd_normalize_tiles(const void* slots, int n_src, int n_freq_local, int n_chan, float pre_scale,
                             int out_kind, void* out, void* stream);
/*
 * Several ranks, all rows of a rank in one launch: the exchange fused into the contraction on both sides.
 * spyb_csd_accumulate_tiles_others is spyb_csd_accumulate_tiles without the frequencies rank src_rank owns itself;
 * after the barrier spyb_csd_coherence_planar_slots runs those (planes / n_freq = the local slab) like
 * spyb_csd_coherence_planar, its epilogue adding the tiles the peers stored into the local slot buffer
 * slots [n_src][n_freq][n_tiles][128][128] (source skip_src = this rank is not read) before it normalises: the
 * trial sum over ranks (computational_routine.py:1022-1032, the reference's locked HDF5 "+=",
 * shared/kwarg_decorators.py:722-735) and csd.py:118-172 without a separate reduction or normalisation kernel.
 * The peers must have used alpha = 1 (coherency does not depend on a common factor).
 */
int spyb_csd_accumulate_tiles_others(const float* planes, long long sx_f, long long sx_r, int n_rows, int n_freq,
                                     int n_chan, float alpha, float beta, void* const* owner_base_host,
                                     const int* f_begin_host, int n_owners, int src_rank, void* stream);
int spyb_csd_coherence_planar_slots(const float* planes, long long sx_f, long long sx_r, int n_rows, int n_freq,
                                    int n_chan, const void* slots, int n_src, int skip_src, int out_kind, void* out,
                                    void* stream);

/*
 * Slot buffers that other ranks of the node can write into (one process per GPU): plain cudaMalloc memory plus
 * its 64-byte CUDA IPC handle; a peer rank maps it with spyb_peer_open (which also enables P2P access between
 * its device and the owner's, NVLink on an NVSwitch node) and passes the mapped pointer as owner_base_host[o].
 */
int spyb_peer_alloc(long long bytes, void** ptr_out, unsigned char* handle64_out);
int spyb_peer_open(const unsigned char* handle64, void** ptr_out);
/* stream-ordered byte fill of local or peer-mapped slot memory (a rank without trials clears its source slots: its
 * peers add whatever those slots hold) */
int spyb_peer_memset(void* ptr, int value, long long bytes, void* stream);
int spyb_peer_close(void* mapped_ptr);
int spyb_peer_free(void* ptr);

/*
 * Coherency + output conversion:  out = conv( pre*C_ij / sqrt(pre*C_ii * pre*C_jj) ).
 * Replaces: syncopy/connectivity/csd.py:118-172 (`normalize_csd`).
 *   csd [n_mat][n_chan][n_chan] complex64; out float32 or complex64 of the same shape.
 */
int spyb_csd_normalize(const void* csd, long long n_mat, int n_chan, float pre_scale, int out_kind,
                       void* out, void* stream);

/*
 * Whole-trial detrend: out = x - (mean [+ slope * t]) per channel, float32 (scipy.signal.detrend along time as
 * called at syncopy/specest/compRoutines.py:582-585, 751-753).  polyremoval: -1 copy, 0 de-mean, 1 linear.
 */
int spyb_detrend(const float* x, int n_trials, long long trial_stride, int n_samples, int n_chan, int polyremoval,
                 float* out, long long out_trial_stride, void* stream);

/*
 * Wavelet / superlet transform as FFT convolution with per-scale products of powers:
 *     z[trial][n][s][c] = prod_{j < n_fac[s]} ( conv_same(x[trial][:, c], psi_{s,j})[n] ) ^ expo[s][j]
 * Replaces: syncopy/specest/wavelets/transform.py:88-108 (`cwt_time`; one factor, exponent 1),
 *           syncopy/specest/superlet.py:108-198, 321-365 (`multiplicativeSLT`, `FASLT`, `cwtSL`) and the output
 *           conversion of syncopy/specest/compRoutines.py:593-595, 760-762.
 *   xspec   complex64 [n_trials][n_dft/2+1][n_chan]: one-sided FFT_{n_dft} of the zero-padded (detrended) trials,
 *           as spyb_mtmfft writes it with a unit taper, scale 1 and nfft = n_dft
 *   kern    complex64 [n_scales][max_fac][n_dft] = FFT_{n_dft}(h_{s,j}) / n_dft with h[d mod n_dft] = psi[d + (M-1)/2]
 *           (psi sampled by the caller in float64 exactly as the reference does; n_dft a power of two <= 16384 or
 *           any even length with prime factors <= 61 beyond that, >= n_samples + the kernel's half support)
 *   expo    float32 [n_scales][max_fac] exponents (principal-branch complex power; 1 = none)
 *   n_fac   int32 [n_scales] number of factors in use
 *   out     [n_trials][n_time][n_scales][n_chan] float32 or complex64 by out_kind; rows n = 0 .. n_time-1
 *   transposed  1: xspec is [n_trials][n_chan][n_dft/2+1] and out is [n_trials][n_scales][n_chan][n_time], the
 *           layouts in which every global access of the kernel is contiguous across a warp (a block owns one
 *           channel of one scale); spyb_transpose converts to / from the reference layouts at HBM speed
 */
int spyb_cwt(const void* xspec, int n_trials, int n_chan, int n_dft, const void* kern, const float* expo,
             const int* n_fac, int n_scales, int max_fac, int n_time, int out_kind, int transposed, void* out,
             void* stream);

/* batched 2-D transpose of 4- or 8-byte elements: in [batch][rows][cols] -> out [batch][cols][rows] */
int spyb_transpose(const void* in, void* out, int batch, int rows, int cols, int elem_bytes, void* stream);

/*
 * Transpose with a placement: in [nb1 * nb2][rows][cols] -> out element (b1, b2, col, row) at
 * b1*stride_b1 + b2*stride_b2 + col*ld_out + row (element units); columns of segment b2 are kept while
 * b2*cols + col < col_limit.  Lands the time-contiguous rows of a wavelet launch over a range of scales -- or over
 * overlap-save segments of the trial, the "trials" of that launch being (trial, segment) pairs -- in its slice of
 * the result [n_trials][n_time][n_scales][n_chan] (transform.py:88-108: one fftconvolve per scale; short kernels do
 * not need the full padded length).
 */
int spyb_transpose_place(const void* in, void* out, int nb1, int nb2, int rows, int cols, int elem_bytes,
                         long long stride_b1, long long stride_b2, long long ld_out, int col_limit, void* stream);

/* dst[t][i][:] = src[t][idx[i]][:], rows of row_elems float32 (time post-selection, compRoutines.py:593) */
int spyb_gather_rows(const float* src, int n_trials, long long src_trial_stride, const int* idx, int n_idx,
                     long long row_elems, float* dst, void* stream);

/*
 * In place on complex64 [n_freq][n_chan][n_chan]: lower triangle <- conjugate of the upper one, diagonal made real.
 * Restores exact Hermitian symmetry of a trial-summed cross-spectral matrix after a reduction over ranks whose
 * addition order differs between (i, j) and (j, i) (the reference sums trial by trial in one process,
 * computational_routine.py:1022-1032, so both halves see the same order); wilson_sf.py:104-106 measures the
 * factorisation error element-wise and never gets below such an asymmetry.
 */
int spyb_csd_mirror_upper(void* csd, int n_freq, int n_chan, void* stream);

/* x *= s on n float32 (trial mean `/= nTrials`, computational_routine.py:1030-1032) */
int spyb_scale(float* x, long long n, float s, void* stream);

/*
 * acc[e] = beta*acc[e] + alpha * sum_{b < n_trials} src[b*trial_stride + e], e < n_elems (float32; complex64 data counts
 * two floats per element).  Replaces the runtime's trial accumulation `target[()] += res` and the final
 * `target[()] /= numTrials` of syncopy/shared/computational_routine.py:1025,1030-1032 for a batch of trials.
 * n_elems, trial_stride multiples of 4, pointers 16-byte aligned.
 */
int spyb_sum_trials(const float* src, int n_trials, long long trial_stride, long long n_elems, float alpha, float beta,
                    float* acc, void* stream);

/*
 * Jackknife over trials (syncopy/statistics/jackknifing.py:14-184) -- element-wise pieces on float32 words
 * (a complex64 array counts two words per element):
 *   spyb_axpby            out = a*x + b*y (y may be NULL): leave-one-out replicate (T*avg - x_k)/(T-1) (:80-85) and
 *                         bias = (T-1)*(jack_avg - direct) (:160)
 *   spyb_sqdev_accumulate var[e] += |avg[e] - x[e]|^2 over n_elem elements, real or complex input (:164-170)
 */
int spyb_axpby(const float* x, const float* y, float a, float b, float* out, long long n, void* stream);
int spyb_sqdev_accumulate(const float* avg, const float* x, float* var, long long n_elem, int is_complex, void* stream);

/*
 * Pairwise phase consistency (syncopy/connectivity/ST_compRoutines.py:158-233 `ppc_column_cF` and the pair loop
 * of connectivity_analysis.py:624-667): the average of cos(angle(z_j conj z_k)) over all trial pairs j < k equals
 * (|sum_k z_k/|z_k||^2 - T) / (T (T - 1)), so one pass over the T single-trial cross spectra replaces T(T-1)/2 pair
 * evaluations.  spyb_unit_accumulate adds the unit vectors of one trial's cross spectra (complex64, n elements;
 * `first` != 0 initialises acc), spyb_ppc_finish turns the sums into the float32 PPC.
 */
int spyb_unit_accumulate(const void* z, void* acc, long long n, int first, void* stream);
int spyb_ppc_finish(const void* acc, float* out, long long n, int n_trials, void* stream);

/*
 * Cross-covariance of one trial (syncopy/connectivity/ST_compRoutines.py:465-584): the C^2 `fftconvolve(x_i,
 * x_j[::-1], 'same')` calls become one forward spectrum per channel (spyb_mtmfft, unit taper, n_dft >= 2 n - 1), kernel
 * spectra of the reversed channels (spyb_xcov_kernel_spectra; xspec [chan][n_dft/2+1] complex64 ->
 * kern [chan][n_dft] complex64, circular result advanced by `shift` samples), the batched inverse transforms of
 * spyb_cwt (real output, n_time = 2 n_lags + 1, shift = n - 1 - n_lags) and spyb_xcov_finish, which picks the lags
 * (incl. the one-sample asymmetry of 'same' for even n, :555-566), divides by the overlap n - s and optionally by the
 * channel standard deviations (:569-572).  out [n_lags][C][C] float32.
 */
int spyb_xcov_kernel_spectra(const void* xspec, int n_chan, int n_dft, int n_samples, int shift, void* kern, void* stream);
int spyb_xcov_finish(const float* corr, const void* xspec, int n_chan, int n_samples, int n_lags, int n_dft, int norm,
                     float* out, void* stream);

/*
 * Preprocessing compute functions (syncopy/preproc/compRoutines.py).
 *   spyb_sosfilt      IIR filtering with second-order sections, replaces scipy.signal.sosfilt (twopass = 0: zero initial
 *                     state) and scipy.signal.sosfiltfilt (twopass = 1: odd extension by `edge` samples, steady-state
 *                     initial conditions `zi_host` [n_sections][2] scaled by the edge sample, forward + backward pass) as
 *                     called by but_filtering_cF (:175-276).  sos_host [n_sections][6] float64 (host), recursion in
 *                     float64, out float32 [trial][sample][channel]; scratch float64 [trial][n_samples + 2 edge][channel].
 *   spyb_upfirdn      polyphase FIR resampling, replaces scipy.signal.upfirdn inside resample_poly as called by
 *                     resample_cF (:541-616, preproc/resampling.py:14-79): out[m] = sum_i x[i] h[(m + first_row) down -
 *                     i up], h float64 [len_h] (device; already zero-padded and scaled by `up` like resample_poly does),
 *                     accumulation in float64.
 *   spyb_standardize  (x - mean) / std per channel (standardize_cF, :765-832; detrending first via spyb_detrend).
 *   spyb_rectify      |x| (rectify_cF, :303-338).
 * The windowed-sinc FIR filters of sinc_filtering_cF (:27-148, preproc/firws.py) and the analytic signal of hilbert_cF
 * (:365-419) are 'same' / circular convolutions and run on spyb_cwt with host-designed kernels.
 */
int spyb_sosfilt(const float* x, int n_trials, long long trial_stride, int n_samples, int n_chan, const double* sos_host,
                 int n_sections, const double* zi_host, int edge, int twopass, double* scratch, float* out, void* stream);
int spyb_upfirdn(const float* x, int n_trials, long long trial_stride, int n_in, int n_chan, const double* h, int len_h,
                 int up, int down, int first_row, int n_out, float* out, void* stream);
int spyb_standardize(const float* x, int n_trials, long long trial_stride, int n_samples, int n_chan, float* out, void* stream);
int spyb_rectify(const float* x, float* out, long long n, void* stream);

/*
 * Granger causality path (float64 / complex128 like the reference, AV_compRoutines.py:395).  These three calls
 * synchronise `stream` internally: the regularisation ladder and Wilson's iteration are data dependent.
 * `work` is caller-owned device scratch of at least spyb_*_workspace_bytes().  `_host` pointers are host memory.
 *
 * spyb_regularize_csd replaces syncopy/connectivity/wilson_sf.py:197-254 (`regularize_csd`):
 *   csd       [n_freq][n_chan][n_chan] complex64 (trial-averaged cross spectra)
 *   out       [n_freq][n_chan][n_chan] complex128 = csd + eps*I (eps = 0 when no regularisation was needed)
 *   eps_host  0 (not needed), the factor used, or -1 (cond_max not reached even with eps_max; out then holds the
 *             last attempt, csd + eps_max*I)
 *   cond0_host  largest 2-norm condition number over the frequencies of the input
 *   The condition numbers come from the extreme |eigenvalues| of the Hermitian matrices (Householder
 *   tridiagonalisation + Sturm bisection in float64) instead of LAPACK's SVD.
 *
 * spyb_wilson replaces syncopy/connectivity/wilson_sf.py:16-194 (`wilson_sf`, direct_inversion=True):
 *   csd       [n_freq][n_chan][n_chan] complex128, one-sided spectrum (the mirror to negative frequencies of
 *             wilson_sf.py:63 is implicit), positive definite; n_chan <= 256
 *   H         [n_freq][n_chan][n_chan] complex128 transfer function, Sigma [n_chan][n_chan] float64 noise covariance
 *   converged_host / err_host / iters_host   convergence flag, final max relative error, iterations done
 *   Returns non-zero (message: "not positive definite" / "singular") where NumPy would raise LinAlgError.
 *
 * spyb_granger replaces syncopy/connectivity/granger.py:10-79:
 *   out [n_freq][n_chan][n_chan] float32, out[f][i][j] = Granger causality i -> j.
 */
long long spyb_regularize_workspace_bytes(int n_freq, int n_chan);
int spyb_regularize_csd(const void* csd, int n_freq, int n_chan, double cond_max, double eps_max, int n_steps,
                        void* out, double* eps_host, double* cond0_host, void* work, long long work_bytes,
                        void* stream);
long long spyb_wilson_workspace_bytes(int n_freq, int n_chan);
int spyb_wilson(const void* csd, int n_freq, int n_chan, int n_iter, double rtol, void* H, double* Sigma,
                int* converged_host, double* err_host, int* iters_host, void* work, long long work_bytes,
                void* stream);
int spyb_granger(const void* csd, const void* H, const double* Sigma, int n_freq, int n_chan, float* out,
                 void* stream);

/*
 * spyb_wilson over several ranks that all hold the same (all-reduced) csd: everything in an iteration except the
 * plus operator (wilson_sf.py:154-184) is independent per frequency, so rank r only factorises its slab
 * [f_lo, f_hi); the plus operator needs every frequency of every matrix element and is replicated.  The library
 * calls `exchange` twice per iteration, in stream order:
 *   what = 0: buf = lag-domain work array, n_rows = 2(n_freq-1) rows of row_bytes bytes; the rank has written rows
 *             [f_lo, f_hi) and the mirror rows 2(n_freq-1) - f for f in [max(f_lo,1), min(f_hi, n_freq-1));
 *             the callback must bring in the rows of all other ranks (e.g. one NCCL broadcast per rank and range);
 *   what = 1: buf = one float64, replace it by its maximum over the ranks.
 * A non-zero return of the callback aborts.  H is written for the slab only; Sigma, converged, err, iterations are
 * identical on every rank.  `exchange_ctx` is passed through.
 */
typedef int (*spyb_exchange_fn)(void* ctx, int what, void* buf, long long row_bytes, int n_rows);
int spyb_wilson_sharded(const void* csd, int n_freq, int n_chan, int n_iter, double rtol, void* H, double* Sigma,
                        int* converged_host, double* err_host, int* iters_host, void* work, long long work_bytes,
                        int f_lo, int f_hi, spyb_exchange_fn exchange, void* exchange_ctx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPYB200_H */
