/*
 * epb200.h - C ABI of libepb200.so: the B200 (sm_100a) implementation of echopype's
 * calibrate -> clean -> commongrid array-compute path.
 *
 * echopype (pure Python) has no FFI of its own; each entry point below replaces one internal seam of
 * the reference where a whole (channel, ping_time, range_sample) array changes hands.  The reference
 * interface each function stands in for is cited as file:line under /root/reference/echopype.
 * INTEGRATION.md shows the ctypes stub a maintainer would add on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name starts with h_; the caller owns all memory;
 *     the library never allocates or frees; scratch sizes are returned by epb_*_workspace_bytes;
 *   - all calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream)
 *     of the CURRENT device and never synchronise;
 *   - arrays are C-contiguous in the reference's dimension order (channel, ping_time, range_sample
 *     [, beam]); C, P, R, B are the sizes of those dimensions;
 *   - return value 0 = ok, negative = EPB_E_* ; a message is available from epb_last_error()
 *     (thread-local); no C++ exception crosses the boundary;
 *   - NaN (quiet) marks invalid samples everywhere, as in the reference.
 */
#ifndef EPB200_H
#define EPB200_H

#ifdef __cplusplus
extern "C" {
#endif

#define EPB_VERSION 100 /* 0.1.0 */

#define EPB_OK 0
#define EPB_E_BADARG (-1)   /* NULL pointer, non-positive size, unsupported option            */
#define EPB_E_CUDA (-2)     /* kernel launch / runtime error (cudaGetLastError text available) */
#define EPB_E_UNSUPPORTED (-3)

typedef long long epb_i64;

/* A (channel, ping_time) float64 parameter on the device with element strides; stride 0 broadcasts
 * (scalar: sc = sp = 0; per-channel: sp = 0). */
typedef struct epb_cp {
  const double* ptr;
  epb_i64 sc;
  epb_i64 sp;
} epb_cp;

/* Per-(channel, ping) row record (192 bytes) consumed by every sample kernel.  It carries
 *   (1) the EXACT float64 range law, evaluated with the reference's operation order, used only for
 *       index-space decisions (first sample with R' > 0, bin boundaries) so that those decisions are
 *       bit-identical with the float64 reference:
 *         law 0 (EK,  calibrate/range.py:138):   R(n) = ((n * p0) * p1) / 2          p0 = sample_interval, p1 = sound_speed
 *         law 1 (AZFP, calibrate/range.py:81-89): R(n) = (p0 + p1 * (k / p2 + p3)) - p4,  k = (2(n+1)-1)*N - 1
 *                                                 p0 = c*L/(2f), p1 = c/4, p2 = f, p3 = tau, p4 = TS offset, N = azfp_N
 *   (2) the affine value form R = r0 + a*n and the folded calibration constants:
 *         out = front(x) + slog*log10(R') + two_alpha*R' - K,   R' = (R - off1) - off2
 *         front(x) = x*fscale + foff (power dB / AZFP counts) or 10*log10(fscale*|mean_beam x|^2) (complex)
 *   (3) the same value form pre-split into float32 hi/lo pairs (written by the epb_rows_* kernels) so the
 *       per-sample kernels stay on the FP32 pipes: R' = fma(a_h,n,bp_h) + fma(a_l,n,bp_l) is accurate to
 *       ~1 ulp(R') even where R - off cancels.
 */
typedef struct epb_row {
  double p0, p1, p2, p3, p4;
  double off1, off2;
  double r0, a;
  double two_alpha;
  double K;
  double fscale, foff;
  double slog;
  int n_start; /* first n with R' > 0 under the exact law (R when none); 0 when the variant has no guard */
  int law;     /* 0 EK, 1 AZFP; bit 8 set: echo_range is NaN where the input sample is NaN (EK) */
  int azfp_N;
  int reserved;
  /* float32 block (offset 128) */
  float a_h, a_l;     /* a = a_h + a_l                                   */
  float r0_h, r0_l;   /* r0                                              */
  float bp_h, bp_l;   /* r0 - off1 - off2 (intercept of R')              */
  float two_alpha_f;  /* 2*alpha                                         */
  float slog2;        /* slog*log10(2): slog*log10(R') = slog2*log2(R')  */
  float fscale_f;     /* front scale                                     */
  float foffK;        /* foff - K                                        */
  float c0;           /* (foff - K) * log2(10)/10  : 10^((front(x)-K)/10) = 2^(x*c1 + c0) */
  float c1;           /* fscale * log2(10)/10                             */
  float c2;           /* two_alpha * log2(10)/10                          */
  float spow;         /* slog/10: R'^spow is the linear-domain spreading  */
  double range_last;  /* exact range law at the last sample, law_range(R - 1): upper bound of echo_range in the row */
} epb_row;

#define EPB_LAW_EK 0
#define EPB_LAW_AZFP 1
#define EPB_LAW_NANRANGE 256

#define EPB_CAL_SV 0
#define EPB_CAL_TS 1
#define EPB_SONAR_EX60 0 /* EK60 / ES70: TVG offset = 2 samples (range.py:176-178)                 */
#define EPB_SONAR_EX80 1 /* EK80 / ES80 / EA640: c*tau/4, GPT channels additionally 2 samples (:180-199) */

const char* epb_last_error(void);
int epb_version(void);

/* ---- row setup (replaces the (channel,ping) parameter broadcasting in CalibrateEK._cal_power_samples,
 *      calibrate/calibrate_ek.py:98-183, range_mod_TVG_EK calibrate/range.py:160-201) -------------------- */
int epb_rows_ek_power(epb_row* rows, epb_i64 C, epb_i64 P, epb_i64 R, int sonar, int cal_type,
                      epb_cp sample_interval, epb_cp sound_speed, epb_cp absorption,
                      epb_cp tau_nominal, epb_cp transmit_power, epb_cp gain, epb_cp sa_correction,
                      epb_cp equivalent_beam_angle, epb_cp frequency_nominal, epb_cp tau_effective,
                      const unsigned char* is_gpt /* [C] or NULL */, void* stream);

/* calibrate/calibrate_azfp.py:49-111 + compute_range_AZFP calibrate/range.py:11-95.
 * Per-channel arrays are [C] float64 on the device. */
int epb_rows_azfp(epb_row* rows, epb_i64 C, epb_i64 P, epb_i64 R, int cal_type, epb_cp sound_speed,
                  epb_cp absorption, epb_cp tau_nominal, const double* N, const double* f_dig,
                  const double* L, const double* EL, const double* DS, const double* TVR,
                  const double* VTX0, const double* psi_linear, const double* Sv_offset, void* stream);

/* CalibrateEK80._cal_complex_samples calibrate/calibrate_ek.py:532-659 (constants of :613-637 and the
 * received-power scale of :483-490).  `gain` already includes the BB B_theta_phi term, `psi` the BB
 * frequency scaling; waveform_bb != 0 drops the 2*Sa term. */
int epb_rows_ek80_complex(epb_row* rows, epb_i64 C, epb_i64 P, epb_i64 R, int cal_type, int waveform_bb,
                          int n_beam, epb_cp sample_interval, epb_cp sound_speed, epb_cp absorption,
                          epb_cp tau_nominal, epb_cp transmit_power, epb_cp gain, epb_cp sa_correction,
                          epb_cp psi, epb_cp freq_center, epb_cp tau_effective, epb_cp z_et, epb_cp z_er,
                          const unsigned char* is_gpt, void* stream);

/* ---- K1: fused per-sample Sv/TS for power samples (CalibrateEK._cal_power_samples calibrate_ek.py:79,
 *      CalibrateAZFP._cal_power_samples calibrate_azfp.py:49, compute_range_EK range.py:98) -------------
 * out, echo_range: [C,P,R] float32 (echo_range may be NULL).  minmax: NULL or 4 floats
 * {min(out), max(out), min(range), max(range)} updated atomically (initialise with epb_minmax_init). */
int epb_sv_power(const float* backscatter_r, const epb_row* rows, float* out, float* echo_range,
                 float* minmax, epb_i64 C, epb_i64 P, epb_i64 R, void* stream);

/* K1 on RAW POWER COUNTS: counts [C,P,R] int16 (-32768 = NaN padding) are scaled by INDEX2POWER in
 * registers (the values epb_ingest_power_i16 writes), so the kernel reads 2 bytes per sample.  Needs R % 4 == 0. */
int epb_sv_power_i16(const short* counts, const epb_row* rows, float* out, float* echo_range, float* minmax, epb_i64 C,
                     epb_i64 P, epb_i64 R, void* stream);

/* ---- K2: complex CW samples (CalibrateEK80._get_power_from_complex calibrate_ek.py:456-505 without
 *      pulse compression + _cal_complex_samples).  re/im: [C,P,R,B] float32, B <= 4. ------------------- */
int epb_sv_complex(const float* re, const float* im, const epb_row* rows, float* out, float* echo_range,
                   float* minmax, epb_i64 C, epb_i64 P, epb_i64 R, int B, void* stream);

/* ---- K3: broadband pulse compression + Sv/TS epilogue (compress_pulse calibrate/ek80_complex.py:316,
 *      _convolve_per_channel :285, get_norm_fac :372, then calibrate_ek.py:483-490,:581,:613-637).
 * replica: per-channel conj-free transmit replicas tx[c][k] as float2, concatenated; replica_off[C+1]
 * (host array) gives each channel's start; inv_norm[C] = 1/||tx||^2 (device, float64).
 * pc_out (optional): [C,P,R] float2 beam-averaged normalised pulse-compressed signal. */
int epb_pulse_compress_sv(const float* re, const float* im, const float* replica /* float2 */,
                          const int* h_replica_off, const double* inv_norm, const epb_row* rows,
                          float* out, float* echo_range, float* pc_out, float* minmax, epb_i64 C,
                          epb_i64 P, epb_i64 R, int B, void* stream);
/* K3 by overlap-save FFT (scipy.signal.convolve, which compress_pulse calls at ek80_complex.py:310, is free to choose the
 * FFT method as well): same arguments and results as epb_pulse_compress_sv for replicas of up to
 * epb_pulse_fft_max_taps() taps, ~17x less arithmetic at 277 taps and about half the rounding error of the direct
 * float32 sum.  workspace: caller-owned device scratch of epb_pulse_fft_workspace_bytes(C) bytes (16-byte aligned) that
 * receives the replica spectra and the twiddle tables of this launch. */
epb_i64 epb_pulse_fft_workspace_bytes(epb_i64 C);
int epb_pulse_fft_max_taps(void);
int epb_pulse_compress_sv_fft(const float* re, const float* im, const float* replica /* float2 */,
                              const int* h_replica_off, const double* inv_norm, const epb_row* rows,
                              float* out, float* echo_range, float* pc_out, float* minmax, epb_i64 C,
                              epb_i64 P, epb_i64 R, int B, void* workspace, epb_i64 workspace_bytes, void* stream);

/* ---- K4/K5: De Robertis & Higginbottom background noise (estimate_background_noise clean/api.py:362,
 *      remove_background_noise :436) on materialised Sv / echo_range -------------------------------------
 * noise: [C, ceil(P/ping_num)] float32 (dB).  alpha: (channel,ping) absorption.  noise_max: NaN = no cap. */
int epb_noise_estimate(const float* Sv, const float* echo_range, epb_cp absorption, float* noise,
                       epb_i64 C, epb_i64 P, epb_i64 R, int ping_num, int range_sample_num,
                       float noise_max, void* stream);
/* Sv_noise and Sv_corrected may each be NULL.  snr_threshold in dB. */
int epb_noise_apply(const float* Sv, const float* echo_range, epb_cp absorption, const float* noise,
                    float* Sv_noise, float* Sv_corrected, float* minmax /* 4 floats or NULL */,
                    epb_i64 C, epb_i64 P, epb_i64 R, int ping_num, float snr_threshold, void* stream);

/* ---- K6: linear-domain bin reduction for MVBS / NASC (_groupby_x_along_channels
 *      commongrid/utils.py:504-628 = flox xarray_reduce; compute_raw_NASC :97-207) ----------------------
 * xbin: [P] int32 bin of each ping along ping_time / distance (-1 = outside every bin).
 * Generic form: range values given per sample (float32 or float64), edges [nR+1] float64 (ascending).
 * acc: [C, nX, nR, 4] float64 accumulators {sum of 10^(Sv/10) over non-NaN members, #non-NaN members,
 * #NaN members, sum of range differences r[n+1]-r[n] of the members (NASC heights, with_height != 0)};
 * zero it with epb_zero before the first call; several calls (and several ranks) may accumulate. */
int epb_bin_reduce(const float* Sv, const void* range_var, int range_is_f64, const int* xbin,
                   const double* r_edges, int nR, int closed_right, int with_height, double* acc,
                   epb_i64 C, epb_i64 P, epb_i64 R, epb_i64 nX, void* stream);
/* Same reduction when the range variable was produced by this library from `rows` (echo_range of
 * compute_Sv, or depth = depth_off[p] + echo_range*depth_scale[p], consolidate/api.py:221): bin
 * boundaries are located in sample-index space by bisection on the exact float64 range law, so bin
 * membership is bit-identical with binning the reference's float64 echo_range.  A NaN input sample is
 * counted as a NaN member (the reference drops it because its echo_range is NaN): identical results for
 * skipna=True with a NaN fill_value; use epb_bin_reduce otherwise.  nR <= 511. */
int epb_bin_reduce_law(const float* Sv, const epb_row* rows, const double* depth_off /* [P] or NULL */,
                       const double* depth_scale /* [P] or NULL */, const int* xbin, const double* r_edges,
                       int nR, int closed_right, double* acc, epb_i64 C, epb_i64 P, epb_i64 R, epb_i64 nX,
                       void* workspace /* NULL or epb_pipeline_workspace_bytes(C, P, 0) bytes, 16-byte aligned: regular
                       echo_range volumes then run on the persistent kernel of the fused pipeline */,
                       epb_i64 workspace_bytes, void* stream);
/* mean -> dB.  out: [C,nX,nR] float32.  skipna=0 reproduces func="mean".  Bins without members get
 * fill_value (then 10log10 like the reference).  h_out (NASC, optional): sum of heights per bin. */
int epb_bin_finalize(const double* acc, float* out, double* h_out, epb_i64 ncell, int skipna,
                     float fill_value, int to_db, void* stream);

/* ---- K7: index binning (compute_MVBS_index_binning commongrid/api.py:195-266: coarsen mean in the
 *      linear domain + coarsen min of echo_range).  out/er_out: [C, ceil(P/pn), ceil(R/rn)] float32. ---- */
int epb_coarsen(const float* Sv, const float* echo_range, float* out, float* er_out, epb_i64 C,
                epb_i64 P, epb_i64 R, int ping_num, int range_sample_num, void* stream);

/* ---- fused pipeline: power -> Sv -> background-noise removal -> MVBS accumulators in one pass over
 *      HBM (compute_Sv -> remove_background_noise -> compute_MVBS).
 * rows: Sv rows from epb_rows_ek_power / epb_rows_azfp.  r_edges [nR+1] float64 (range bins on echo_range),
 * xbin [P], acc as in epb_bin_reduce.  Optional full-size outputs (any may be NULL): Sv, echo_range, Sv_noise,
 * Sv_corrected [C,P,R] float32.  noise_out: [C, ceil(P/ping_num)] (dB) or NULL.
 * ping_num = 0 skips noise removal (Sv -> MVBS).  noise_max: dB, NaN = no cap.  Requires R % 4 == 0,
 * ping_num <= 256; returns EPB_E_UNSUPPORTED when the per-CTA accumulators exceed shared memory.
 * range_max_out: NULL or one float64 that receives the exact nanmax(echo_range) (what epb_range_max computes; -inf if
 * every range is NaN), for trimming an upper-bound range grid (commongrid/api.py:108-115) without a second pass.
 * workspace: NULL or a 16-byte aligned device scratch buffer of workspace_bytes >=
 * epb_pipeline_workspace_bytes(C, P, ping_num) bytes owned by the caller and private to this launch (it receives
 * one 144-byte descriptor per ping tile); with a workspace, regular volumes (every ping tile shares one range law, finite
 * calibration constants, R <= 8192 (float32 input; full-size outputs and int16 counts: R <= 4096), ping_num <= 64) run on the persistent
 * register-resident kernel (pipeline_fast_impl.cuh), decided on the device without a host synchronisation. */
int epb_pipeline_power_mvbs(const float* backscatter_r, const epb_row* rows, const int* xbin,
                            const double* r_edges, int nR, int closed_right, double* acc, float* noise_out,
                            float* Sv, float* echo_range, float* Sv_noise, float* Sv_corrected, epb_i64 C,
                            epb_i64 P, epb_i64 R, epb_i64 nX, int ping_num, int range_sample_num,
                            float noise_max, float snr_threshold, double* range_max_out, void* workspace,
                            epb_i64 workspace_bytes, void* stream);
/* The same pipeline on RAW POWER COUNTS: counts [C,P,R] int16 as read from the power datagrams,
 * -32768 marking the NaN padding of shorter pings (convert/parse_base.py:686-730); the samples enter the kernel at
 * 2 bytes each and are scaled by INDEX2POWER (parse_base.py:24,302) in registers, bit-identical with
 * epb_ingest_power_i16 followed by epb_pipeline_power_mvbs.  scratch: [C,P,R] float32 owned by the caller; it is
 * written (by the ingest kernel) only when the volume is irregular and the general kernel has to run.
 * Needs R % 8 == 0 for the direct path; no full-size outputs (ingest + epb_pipeline_power_mvbs serve those). */
int epb_pipeline_power_mvbs_i16(const short* counts, float* scratch, const epb_row* rows, const int* xbin,
                                const double* r_edges, int nR, int closed_right, double* acc, float* noise_out,
                                epb_i64 C, epb_i64 P, epb_i64 R, epb_i64 nX, int ping_num, int range_sample_num,
                                float noise_max, float snr_threshold, double* range_max_out, void* workspace,
                                epb_i64 workspace_bytes, void* stream);
epb_i64 epb_pipeline_workspace_bytes(epb_i64 C, epb_i64 P, int ping_num);
/* The same for a given row length: rows of 4097 .. 8192 samples run on the persistent kernel with single-row tiles (one
 * descriptor per ping), which needs a larger workspace than epb_pipeline_workspace_bytes (R <= 4096) reports; with the
 * smaller workspace such volumes take the general kernel. */
epb_i64 epb_pipeline_workspace_bytes_r(epb_i64 C, epb_i64 P, epb_i64 R, int ping_num);
epb_i64 epb_pipeline_smem_bytes(epb_i64 R, int nR, int tile, int do_noise, int staged);
/* SMs the persistent fused kernel leaves free from now on (process-wide, default 0): room for kernels of other streams
 * (the NCCL neighbour exchange of the previous step of a ping-sharded job) next to the persistent grid. */
int epb_set_grid_reserve(int sms);

/* ---- ping-sharded execution: pack / unpack around the ONE all-reduce(sum) that merges the ping bins two
 *      ranks share.  acc [C,nXl,nR,4] float64 (this rank's window of the global ping bins); buf [world][2*C*nR*4 + 1]
 *      float64: own slot = first bin, last bin (zeros when has_last == 0, i.e. nXl == 1) and max(rmax[0..nrmax)), other
 *      slots zero.  After the caller's all-reduce, unpack writes into the first / last local bin the sum of the slices
 *      listed in src [2][world] (entry 2 r = first bin of rank r, 2 r + 1 = its last bin; nsrc0 / nsrc1 entries) and
 *      the maximum of the range maxima into rmax_out (nullable). */
int epb_straddle_pack(const double* acc, epb_i64 C, epb_i64 nXl, epb_i64 nR, const double* rmax, int nrmax, double* buf,
                      int rank, int world, int has_last, void* stream);
int epb_straddle_unpack(const double* buf, const int* src, int nsrc0, int nsrc1, epb_i64 C, epb_i64 nXl, epb_i64 nR,
                        int world, double* acc, double* rmax_out, void* stream);

/* ---- clean.mask_impulse_noise / mask_transient_noise with use_index_binning=True (clean/api.py:30-266; SURVEY.md 8f
 *      rank 3).  nsamp [C] int32: range samples per depth bin of each channel,
 *      ceil(depth_bin / nanmean(diff(range_var))) (clean/utils.py:131-133, 280-282); epb_range_diff_mean returns the
 *      per-channel sum [C] float64 and count [C] uint64 of the valid forward differences of a [C,P,R] float32 range
 *      variable.  epb_first_not_le: first flat index with !(a[i] <= threshold) (np.argmin of the <= mask,
 *      clean/utils.py:141), UINT64_MAX if none. */
int epb_range_diff_mean(const float* range_var, double* sum, unsigned long long* count, epb_i64 C, epb_i64 P, epb_i64 R,
                        void* stream);
int epb_first_not_le(const float* a, epb_i64 n, float threshold, unsigned long long* out, void* stream);
/* Impulse noise (clean/utils.py:263-337): block_means [C,P,nbmax] float32 scratch/output (dB mean of every block of
 * nsamp[c] range samples, NaN-aware, linear domain), mask [C,P,R] uint8 = both ping-wise differences of the forward
 * filled block means (p vs p + k and p vs p - k, NaN difference = +inf) exceed threshold. */
int epb_impulse_noise_mask(const float* Sv, const int* nsamp, float* block_means, unsigned char* mask, epb_i64 C, epb_i64 P,
                           epb_i64 R, int nbmax, int num_side_pings, float threshold, void* stream);
/* Impulse noise with depth-VALUE binning (use_index_binning=False, clean/utils.py:192-260): edges [nbins+1] float64 =
 * np.arange(min, max + depth_bin, depth_bin) of the range variable; per (channel, ping) the NaN-aware linear mean of Sv
 * over the samples of each interval [e_b, e_b+1) (bin_means [C,P,nbins] float32, dB) and the first sample at or below
 * e_b (bin_first [C,P,nbins] int32); every sample takes the mean of its interval (np.digitize + forward fill: upsampled
 * [C,P,R] float32, the reference's upsampled_Sv) and the mask is the two-sided ping comparison of those values.  depth
 * must increase along range_sample (NaN tails allowed).  thresholds_scratch: [nbins+1] float32.
 * upsampled may be NULL: the array is then not materialised and the call is ONE pass over Sv and depth (9 bytes per sample
 * instead of ~25; needs range_sample % 16 == 0, range_sample <= 4096, 16-byte aligned arrays). */
int epb_impulse_noise_mask_depth(const float* Sv, const float* depth, const double* edges, int nbins, float* bin_means,
                                 int* bin_first, float* upsampled, unsigned char* mask, epb_i64 C, epb_i64 P, epb_i64 R,
                                 int num_side_pings, float threshold, float* thresholds_scratch, void* stream);
/* Transient noise (clean/utils.py:109-189 with func = nanmean, clean/api.py:163-166): pooled Sv = dB of the nanmean of
 * 10^(Sv/10) over (2 k + 1) pings x (2 nsamp[c] + 1) range samples of the volume sliced at min_range_sample, borders
 * reflected (scipy.ndimage "reflect"); mask [C,P,R] uint8 = Sv - pooled > threshold, 0 above min_range_sample.
 * window_sums: [C,P,R] float2 scratch, needed only for rows longer than 4096 samples or R % 8 != 0 (may be NULL
 * otherwise: such rows take a single-pass strip kernel); pooled_Sv: NULL or [C,P,R] float32 (NaN above min_range_sample).
 * max_nsamp = max(nsamp) (validated against the sliced axis length: one reflection per border). */
int epb_transient_noise_mask(const float* Sv, const int* nsamp, float* window_sums, unsigned char* mask, float* pooled_Sv,
                             epb_i64 C, epb_i64 P, epb_i64 R, int min_range_sample, int max_nsamp, int num_side_pings,
                             float threshold, void* stream);
/* func = "nanmedian" (clean/api.py:132-145): pooled Sv = dB of np.nanmedian of 10^(Sv/10) over the same windows.  A
 * selection per sample (radix select over the window, O(64 x window)), no scratch. */
int epb_transient_noise_mask_median(const float* Sv, const int* nsamp, unsigned char* mask, float* pooled_Sv, epb_i64 C,
                                    epb_i64 P, epb_i64 R, int min_range_sample, int max_nsamp, int num_side_pings,
                                    float threshold, void* stream);
int epb_transient_noise_mask_depth_median(const float* Sv, const float* depth, unsigned char* mask, float* pooled_Sv,
                                          epb_i64 C, epb_i64 P, epb_i64 R, double depth_min, double depth_max,
                                          double depth_bin, double exclude_above, int num_side_pings, float threshold,
                                          void* stream);

/* Attenuated-signal mask (clean/api.py:269-359 mask_attenuated_signal; clean/utils.py:337-377
 * echopy_attenuated_signal_mask): per (channel, ping) up / lw = np.argmin |range_var - upper_limit_sl| / |range_var -
 * lower_limit_sl| (first minimum, a NaN wins; written to limits [C,P,2] int32), the ping is masked - mask [C,P,R] uint8 =
 * 1 over the WHOLE ping - when 10 log10 nanmedian(10^(Sv[p, up:lw]/10)) - 10 log10 nanmedian(10^(Sv[p-n : p+n, up:lw]/10))
 * < threshold; never within n = num_side_pings pings of either end of the ping axis, nor when Sv[p, up:lw] is empty or
 * all NaN.  The medians are radix selections on the float32 values; the two middle values are averaged in the linear
 * domain in float64.  The "limits outside the range extent" early return (clean/api.py:330-334) is the caller's. */
int epb_attenuated_signal_mask(const float* Sv, const float* range_var, int* limits, unsigned char* mask, epb_i64 C,
                               epb_i64 P, epb_i64 R, double upper_limit_sl, double lower_limit_sl, int num_side_pings,
                               double threshold, void* stream);

/* Transient noise with depth-VALUE windows (use_index_binning=False, clean/utils.py:28-105 pool_Sv, func = nanmean): for
 * every sample whose depth d keeps [d - depth_bin, d + depth_bin] inside [depth_min, depth_max] (the extent of the range
 * variable) and below exclude_above, and whose ping keeps p - k >= 0 and p + k <= P: pooled Sv = dB of the nanmean of
 * 10^(Sv/10) over the samples of the channel with |depth - d| <= depth_bin in pings p - k .. p + k; NaN elsewhere;
 * mask = Sv - pooled > threshold.  depth must increase along range_sample (NaN tails allowed).
 * prefix_sums [C,P,R+1] float64 and prefix_counts [C,P,R+1] int32 are scratch; pooled_Sv NULL or [C,P,R] float32. */
int epb_transient_noise_mask_depth(const float* Sv, const float* depth, double* prefix_sums, int* prefix_counts,
                                   unsigned char* mask, float* pooled_Sv, epb_i64 C, epb_i64 P, epb_i64 R, double depth_min,
                                   double depth_max, double depth_bin, double exclude_above, int num_side_pings, float threshold,
                                   void* stream);
/* The same result for volumes whose depth rows are UNIFORM per channel: every ping has the same depth at a column wherever
 * its depth is defined (a fixed transducer depth and one sample interval / sound speed per channel - the usual case), and
 * samples without a depth (NaN: the padding of shorter pings) carry no Sv either.  The window of a sample is then the same
 * index interval in every ping and the pooled value is a difference of the prefix of the running column sums (the
 * single-pass strip kernel of epb_transient_noise_mask with per-column interval ends): one pass over Sv, no [C,P,R+1]
 * scratch - cfg2: milliseconds instead of 1.6 s.
 * epb_depth_rows_uniform writes the channels' reference depth rows (ref_rows [2,C,R] float32: [0] = column-wise maximum
 * of the defined depths, [1] = scratch) and sets *mismatch (device int) to non-zero when the volume is not uniform (then use
 * epb_transient_noise_mask_depth).  epb_transient_noise_mask_depth_uniform: ref_rows from that call, tables [C,3,R]
 * uint16 scratch; first_column: 0, or a multiple of 16 not beyond the first column of ref_rows that is at or below
 * exclude_above in any channel (shallower columns belong to no window: they are then neither loaded nor scanned);
 * returns EPB_E_UNSUPPORTED unless R % 16 == 0, R - first_column <= 4096 and the arrays are 16-byte aligned. */
int epb_depth_rows_uniform(const float* depth, const float* Sv, float* ref_rows, int* mismatch, epb_i64 C, epb_i64 P, epb_i64 R,
                           void* stream);
int epb_transient_noise_mask_depth_uniform(const float* Sv, const float* depth, const float* ref_rows, unsigned short* tables,
                                           unsigned char* mask, float* pooled_Sv, epb_i64 C, epb_i64 P, epb_i64 R,
                                           double depth_min, double depth_max, double depth_bin, double exclude_above,
                                           int num_side_pings, float threshold, epb_i64 first_column, void* stream);

/* ---- raw power ingest (convert/parse_base.py:24,302 `power = counts.astype(float32) * INDEX2POWER`, :686-730
 *      pad_shorter_ping): n int16 counts -> float32 dB, -32768 (padding marker) -> NaN. -------------------------- */
int epb_ingest_power_i16(const short* counts, float* backscatter_r, epb_i64 n, void* stream);

/* ---- consolidate.add_depth (consolidate/api.py:221): depth = depth_offset[c,p] + echo_range * scale[c,p], scale =
 *      orientation * cos(tilt) (or the platform / beam angle scaling).  echo_range, depth: [C,P,R] float32. -------- */
int epb_add_depth(const float* echo_range, epb_cp depth_offset, epb_cp scale, float* depth, epb_i64 C, epb_i64 P,
                  epb_i64 R, void* stream);

/* ---- mask.frequency_differencing (mask/api.py:593-608): mask[p,n] = (Sv[chanA,p,n] - Sv[chanB,p,n]) op diff with op
 *      0 ">", 1 "<", 2 "<=", 3 ">=", 4 "==" (NaN -> 0); mask: [P,R] uint8.  mask.apply_mask (:437-438): out = mask ?
 *      src : fill_value, mask [C,P,R] (mask_has_channel != 0) or [P,R] broadcast over channels. --------------------- */
int epb_freq_diff_mask(const float* Sv, int chanA, int chanB, int op, float diff, unsigned char* mask, epb_i64 C,
                       epb_i64 P, epb_i64 R, void* stream);
int epb_apply_mask(const float* src, const unsigned char* mask, int mask_has_channel, float fill_value, float* out,
                   epb_i64 C, epb_i64 P, epb_i64 R, void* stream);
/* fill_value given as a (ping_time, range_sample) array, broadcast over channel (mask/api.py:233-246, xr.where :438) */
int epb_apply_mask_fill_array(const float* src, const unsigned char* mask, int mask_has_channel, const float* fill_plane,
                              float* out, epb_i64 C, epb_i64 P, epb_i64 R, void* stream);

/* ---- helpers ---------------------------------------------------------------------------------------- */
int epb_zero(void* ptr, epb_i64 nbytes, void* stream);
int epb_minmax_init(float* minmax /* 4 floats */, void* stream);
/* Global reductions for bin-edge construction (commongrid/api.py:108-114 range_var.max(skipna=True)) and the
 * actual_range attributes (clean/utils.py:392-395).  epb_minmax: minmax[0..1] = min / max over the non-NaN
 * elements of a[0..n), minmax[2] = 1 if any element is NaN (initialise with epb_minmax_init).
 * epb_range_max: exact float64 nanmax of the echo_range that compute_Sv derives from `rows` (NaN where the
 * input sample is NaN; backscatter_r may be NULL when the variant has no such rule); -inf if all NaN. */
int epb_minmax(const float* a, epb_i64 n, float* minmax, void* stream);
int epb_range_max(const float* backscatter_r, const epb_row* rows, epb_i64 C, epb_i64 P, epb_i64 R,
                  double* out_max, void* stream);
/* Philox4x32-10 synthetic inputs: counter = element index / 4, key = seed.
 * kind 0: EK power dB = q * 10log10(2)/256 with q = -24000 + floor(u * 22001 / 2^32) (int16 range)
 * kind 1: AZFP counts = floor(u / 65536) as float
 * kind 2: N(0,1)*scale complex plane (Box-Muller)            nan_tail: fraction of pings (x 2^-16)
 * whose tail beyond a pseudo-random sample index is NaN (pad_shorter_ping). */
int epb_synth_fill(float* out, epb_i64 C, epb_i64 P, epb_i64 R, epb_i64 inner, int kind,
                   unsigned long long seed, epb_i64 ping_offset, unsigned nan_tail_q16, float scale,
                   void* stream);

/* kind 0 of epb_synth_fill as raw int16 counts q (-32768 past the NaN cut): epb_ingest_power_i16 of it equals
 * epb_synth_fill(kind 0) bit for bit. */
int epb_synth_fill_i16(short* out, epb_i64 C, epb_i64 P, epb_i64 R, unsigned long long seed, epb_i64 ping_offset,
                       unsigned nan_tail_q16, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EPB200_H */
