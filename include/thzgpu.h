/* thzgpu.h -- C ABI of libthzgpu, the B200 (sm_100a) implementation of the
 * thz-image-explorer filter-chain hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Each entry
 * point names the reference interface (file:line under the upstream repository) whose
 * arithmetic it replaces.  The reference-side binding a maintainer would add (a Rust
 * `extern "C"` block + `Filter` impls registered with `#[register_filter]`) is shown in
 * INTEGRATION.md.
 *
 * Conventions
 *  - every function returns an int status: THZ_OK (0), THZ_ABORTED (1) or a negative
 *    THZ_E* code; the message is available from thz_last_error().  Nothing throws or
 *    unwinds across this boundary (the reference never propagates errors out of a filter:
 *    src/filters/deconvolution.rs:781-787, 1016-1024).
 *  - cubes are C-contiguous [P][N] f32 with P = width*height and trace index
 *    p = x*height + y, exactly `Array3<f32>` (x, y, t) in standard layout
 *    (src/data_container.rs:109-162; src/math_tools.rs:374 relies on `as_slice()`).
 *    Complex spectra are interleaved (re, im) f32 = `Array3<Complex32>` [P][F], F = N/2+1.
 *  - pointers named d_* are device pointers obtained from thz_dev_alloc(); all others are
 *    host pointers that the library only borrows for the duration of the call.
 *  - there is NO CPU fallback: every compute entry point fails with THZ_ECUDA when no
 *    sm_100 device is usable.
 *  - calls on one context must come from one thread at a time (the reference drives the
 *    chain from a single data thread, src/data_thread.rs:162-174, 1090).
 */
#ifndef THZGPU_H
#define THZGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define THZ_OK 0
#define THZ_ABORTED 1
#define THZ_EINVAL (-1)   /* bad argument / unsupported size */
#define THZ_ECUDA (-2)    /* CUDA runtime error, no device */
#define THZ_ENOMEM (-3)
#define THZ_ESTATE (-4)   /* call order violated (e.g. no plan) */

#define THZ_MAX_BANDS 32
#define THZ_FIR_TAPS 499  /* src/filters/deconvolution.rs:167 */

typedef struct thz_ctx thz_ctx;

/* progress callback, mirrors writes to `progress_lock` (src/filters/deconvolution.rs:774,
 * 896-904, 1026-1037); called on the calling host thread. */
typedef void (*thz_progress_fn)(float fraction, void* user);

/* ---------------------------------------------------------------- context / memory ---- */
/* Replaces the per-struct realfft plans `r2c` / `c2r` (src/data_container.rs:127-129,
 * planned at src/io.rs:614-628): one context per GPU owns streams, twiddle tables, the
 * multiplier vectors and scratch buffers. */
int thz_device_count(void);
int thz_ctx_create(int device, thz_ctx** out);
void thz_ctx_destroy(thz_ctx* ctx);
const char* thz_last_error(const thz_ctx* ctx); /* ctx may be NULL: last create error */
int thz_ctx_device(const thz_ctx* ctx);
int thz_ctx_sm_count(const thz_ctx* ctx);
void* thz_ctx_stream(const thz_ctx* ctx);       /* cudaStream_t all kernels are launched on */
int thz_sync(thz_ctx* ctx);
/* number of kernel launches issued by this context so far (bench.py's gpu_launches) */
int64_t thz_launch_count(const thz_ctx* ctx);

/* FP32 issue-rate microbenchmark on this device (SURVEY.md 8d asks for the fp32 FMA peak measured in the same run
 * as the Richardson-Lucy FLOP fraction).  mode 0 FFMA, 1 packed fma.rn.f32x2, 2 FADD, 3 packed add, 4 FMUL,
 * 5 packed mul; result in lane-operations per second (an FMA counts once: multiply by 2 for FLOP/s). */
int thz_fp32_rate(thz_ctx* ctx, int mode, double* lane_ops_per_s);

int thz_dev_alloc(thz_ctx* ctx, size_t bytes, void** d_ptr);
int thz_dev_free(thz_ctx* ctx, void* d_ptr);
int thz_dev_memset(thz_ctx* ctx, void* d_ptr, int value, size_t bytes);
int thz_copy_h2d(thz_ctx* ctx, void* d_dst, const void* src, size_t bytes);
int thz_copy_d2h(thz_ctx* ctx, void* dst, const void* d_src, size_t bytes);
int thz_host_alloc(size_t bytes, void** ptr);   /* pinned host memory */
int thz_host_free(void* ptr);

/* Synthetic scan: x[p,t] = A_p exp(-((t-t_p)/tau)^2) cos(2 pi f_c (t-t_p))
 * + noise * N(0,1), time[i] = t0 + dt*i, generated on the device with a counter-based RNG.
 * `row0` offsets the pixel index so that row-slab shards generate the same global cube. */
int thz_generate_cube(thz_ctx* ctx, float* d_cube, int width, int height, int n, int row0,
                      int total_width, uint64_t seed, float t0, float dt, float noise);

/* ------------------------------------------------- host-side multiplier vectors -------- */
/* Pixel-independent vectors the reference recomputes for every trace; computed once here, in
 * f32 exactly as the reference does, and handed to thz_plan_trace().  No GPU involved. */
#define THZ_WINDOW_ADAPTED_BLACKMAN 0
#define THZ_WINDOW_BLACKMAN 1
#define THZ_WINDOW_HANNING 2
#define THZ_WINDOW_HAMMING 3
#define THZ_WINDOW_FLAT_TOP 4
/* f[i] = i / (t[n-1] - t[0]), n/2+1 entries (src/io.rs:614-620) */
int thz_frequency_axis(const float* time, int n, float* freq);
/* apply_adapted_blackman_window on a vector of ones (src/math_tools.rs:81-122) */
int thz_adapted_blackman(const float* axis, int n, float lo, float hi, float* mult);
/* the FFT window selected by `FftWindowType` (src/math_tools.rs:356-371, 131-198);
 * lo / hi = ConfigContainer::fft_window, used by the adapted Blackman only */
int thz_window_multiplier(int window_type, const float* time, int n, float lo, float hi, float* mult);
/* TimeDomainBandPassBeforeFFT / AfterFFT (src/filters/band_pass_td_before_fft.rs:134-174): clamps
 * *low / *high to the axis (as the filter mutates its own fields), zero outside [lower, upper),
 * adapted Blackman (ww, ww) inside.  lower / upper may be NULL. */
int thz_time_gate_multiplier(const float* time, int n, double* low, double* high, double window_width,
                             float* mult, int* lower, int* upper);
/* FrequencyDomainBandPass (src/filters/band_pass_fd.rs:135-212) over all f bins */
int thz_band_pass_multiplier(const float* freq, int f, double low, double high, double window_width,
                             float* mult, int* lower, int* upper);

/* -------------------------------------------------------------------- trace plan ------- */
/* The pixel-independent multipliers the chain reduces to (all optional, NULL = ones):
 *   m_pre[n]  : tilt taper x time gate x FFT window
 *               (src/filters/tilt_compensation.rs:188, src/filters/band_pass_td_before_fft.rs:155-174,
 *                src/math_tools.rs:102-198, 356-371)
 *   band[n/2+1]: frequency band-pass taper (src/filters/band_pass_fd.rs:155-212)
 *   m_post[n] : time gate after the inverse FFT (src/filters/band_pass_td_after_fft.rs)
 * n: a power of two in [64, 8192] (radix-16 shared-memory kernels) or any other length in [2, 8192] (chirp-z
 * kernels on the same core, thz_bluestein.cu; above 4096 the 16384-point chirp convolution runs as two
 * 8192-point sub-spectra) -- real scans and tilt-extended axes are not powers of two. */
int thz_plan_trace(thz_ctx* ctx, int n, const float* m_pre, const float* band, const float* m_post);

/* ------------------------------------------------------- trace pass, device pointers --- */
/* Slots 2..7 of the default chain in one kernel: x*m_pre -> rfft -> *band ->
 * irfft/N -> *m_post; img[p] = sum_t out^2 (src/data_thread.rs:1288-1307).
 * d_out may alias d_in.  d_img may be NULL. */
int thz_trace_fused_dev(thz_ctx* ctx, const float* d_in, float* d_out, float* d_img, int64_t P);

/* `math_tools::fft` (src/math_tools.rs:330-398): window (m_pre), unnormalised r2c, |s|,
 * unwrap(arg s) (src/math_tools.rs:211-240).  d_windowed (nullable, may alias d_in) receives
 * the windowed trace the reference leaves in `data` (:356-371). Any output may be NULL. */
int thz_trace_forward_dev(thz_ctx* ctx, const float* d_in, float* d_windowed, float* d_fft,
                          float* d_amp, float* d_phase, int64_t P);

/* `FrequencyDomainBandPass::filter` (src/filters/band_pass_fd.rs:122-220): fft and amplitudes
 * times the plan's band vector (zero outside the band); phases untouched. In place. */
int thz_band_apply_dev(thz_ctx* ctx, float* d_fft, float* d_amp, int64_t P);

/* `math_tools::ifft` (src/math_tools.rs:546-567): unnormalised c2r, / N (imaginary parts of
 * DC and Nyquist ignored).  use_band != 0 multiplies by the plan's band vector first;
 * use_post != 0 multiplies the result by m_post.  d_img (nullable) = sum_t out^2. */
int thz_trace_inverse_dev(thz_ctx* ctx, const float* d_fft, int use_band, int use_post,
                          float* d_out, float* d_img, int64_t P);

/* A time-domain filter that is a pixel-independent multiplier applied on its own stage
 * (`TiltCompensation` at 0 deg, src/filters/tilt_compensation.rs:188; the time gates,
 * src/filters/band_pass_td_before_fft.rs:155-174): out[p][t] = in[p][t] * mult[t].  mult is a
 * host vector of n floats; d_out may alias d_in. */
int thz_time_multiply_dev(thz_ctx* ctx, const float* d_in, const float* mult, int n, float* d_out, int64_t P);

/* `scaling` / `scale_3d` (src/math_tools.rs:242-310): block mean over scale x scale pixels of a
 * [width][height][zlen] f32 array (complex arrays: zlen = 2F); output [width/scale][height/scale][zlen].
 * Same accumulation order as the reference: bit-exact. */
int thz_scale_blocks_dev(thz_ctx* ctx, const float* d_in, int width, int height, int zlen, int scale, float* d_out);
int thz_scale_blocks_host(thz_ctx* ctx, const float* in, int width, int height, int zlen, int scale, float* out);
/* Load path of `open_scan_from_thz` (src/io.rs:578-596): per-trace bias subtraction x <- x - x[0]
 * and the initial intensity image (d_img nullable).  d_out may alias d_in. */
int thz_bias_subtract_dev(thz_ctx* ctx, const float* d_in, int n, float* d_out, float* d_img, int64_t P);

/* `average_polygon_roi` (src/math_tools.rs:599-661): mean over the pixels inside a polygon of a device array
 * [dim0][dim1][zlen] (data, amplitudes or phases); polygon vertices (x, y) are divided by `scaling`; the
 * reference's row flip data[[dim0 - y - 1, x, z]], unsigned ray-casting arithmetic and sequential f32
 * summation order are kept.  `out` is a host vector of zlen floats. */
int thz_roi_average_dev(thz_ctx* ctx, const float* d_data, int dim0, int dim1, int zlen, const int64_t* poly_x,
                        const int64_t* poly_y, int n_points, int scaling, float* out);
/* `calculate_optical_properties` (src/math_tools.rs:663-701), host arithmetic on F values; outputs nullable. */
int thz_optical_properties(const float* sample_amp, const float* sample_phase, const float* ref_amp,
                           const float* ref_phase, const float* freqs, int f, float thickness, float* n_out,
                           float* alpha_out, float* kappa_out);

/* `TiltCompensation` with a non-zero tilt (src/filters/tilt_compensation.rs:97-226).  thz_tilt_plan is the
 * host part (extension steps, extended time axis, per-pixel insert index); thz_tilt_shift_host applies it:
 * out[p][k] = in[p][0] for k < insert[p], in[p][k - insert] * taper[k - insert] up to the end of the trace,
 * 0 after; out is [P][n_ext], n_ext = n + 2 * num_steps. */
int thz_tilt_plan(const float* time, int n, int width, int height, float dx, float dy, double tilt_x, double tilt_y,
                  int* num_steps, float* time_ext, int* insert);
int thz_tilt_shift_host(thz_ctx* ctx, const float* in, const float* taper, const int* insert, int n, int n_ext,
                        float* out, int64_t P);

/* Reference pulse, `ConfigCommand::OpenRef` (src/data_thread.rs:372-588): integer-shift alignment / zero fill of
 * a (ref_time, ref_signal) pulse of m samples to the scan's axis of n samples, FFT window on the reference
 * file's own axis, then r2c, |s| and unwrap(arg s) (the forward kernel on one trace).  signal_out[n],
 * amp_out / phase_out [n/2+1] (nullable).  Replaces the context's trace plan. */
int thz_reference_pulse(thz_ctx* ctx, const float* scan_time, int n, const float* ref_time, const float* ref_signal,
                        int m, int window_type, float window_lo, float window_hi, float* signal_out, float* amp_out,
                        float* phase_out);

/* Voxel opacities of the 3-D view, `instance_from_data` up to the effective threshold
 * (src/gui/threed_plot.rs:165-219; runs on the data thread on a clone of the cube after every chain update):
 * per trace (v^2)^contrast, Gaussian(sigma, radius) weighted sum with zero boundary, traces below
 * opacity_threshold zeroed, the others min/max normalised; *effective_threshold (nullable) = the
 * max_instances-th largest opacity when the cube has more voxels than that, else 0.  The instance list
 * itself (positions, colours) is GUI data and stays on the reference side. */
int thz_voxel_opacity_dev(thz_ctx* ctx, const float* d_cube, int n, int64_t P, float opacity_threshold,
                          float contrast, float sigma, int radius, int64_t max_instances, float* d_opacity,
                          float* effective_threshold);

/* Pixel means that `ifft` computes first (src/math_tools.rs:421-440): mean over all P traces
 * of fft (2F floats), amplitudes (F), phases (F).  Host outputs, any may be NULL. */
int thz_spectral_means(thz_ctx* ctx, const float* d_fft, const float* d_amp, const float* d_phase,
                       int64_t P, float* avg_fft, float* avg_amp, float* avg_phase);

/* ------------------------------------------- reference-normalised spectral maps (config 2) */
/* BASELINE config 2: a scan normalised by a reference pulse (`OpenRef`, src/data_thread.rs:372-588).  The
 * reference program forms A_s / A_r and phi_s - phi_r only inside `calculate_optical_properties` for one pixel or a
 * ROI mean (src/math_tools.rs:665-701, src/data_thread.rs:1489-1559); here the same two operands are produced for
 * every pixel as an epilogue of the forward kernel -- there is no full-map equivalent upstream.
 * thz_plan_reference uploads the reference spectrum (f = n/2 + 1 values each, e.g. from thz_reference_pulse; NULL
 * clears it); thz_trace_forward_normalised_dev is thz_trace_forward_dev whose amplitude output is
 * |s| / max(A_r, 1e-12) and whose phase output is unwrap(arg s) - phi_r; thz_spectral_slice_dev cuts the map of one
 * frequency bin out of a [P][f] array. */
int thz_plan_reference(thz_ctx* ctx, const float* ref_amp, const float* ref_phase, int f);
int thz_trace_forward_normalised_dev(thz_ctx* ctx, const float* d_in, float* d_windowed, float* d_fft, float* d_ratio,
                                     float* d_dphase, int64_t P);
int thz_spectral_slice_dev(thz_ctx* ctx, const float* d_array /* [P][f] */, int f, int bin, float* d_map /* [P] */,
                           int64_t P);

/* --------------------------------------------- GUI hand-off from device-resident cubes -------- */
/* What `data_thread` copies out for the plots after a chain run (src/data_thread.rs:1337-1431), without host copies
 * of the cubes and without spectral cubes: the selected pixel's raw trace (slot 0), its filtered trace (last slot)
 * and its spectrum as slot fft + 1 holds it (windowed r2c; fft and amplitudes times the band-pass, phases
 * untouched) -- one forward transform of one trace with the current plan.  Host outputs, any may be NULL. */
int thz_pixel_handoff_dev(thz_ctx* ctx, const float* d_raw, const float* d_filtered, int64_t P, int64_t pixel,
                          float* raw /* [n] */, float* filtered /* [n] */, float* fft /* [2 f] */, float* amp /* [f] */,
                          float* phase /* [f] */);
/* `avg_signal`: mean over all pixels of a device cube [P][n] (src/data_thread.rs:1423-1431; f64 combination of
 * per-block f32 partial sums instead of the reference's sequential f32 sums). */
int thz_mean_trace_dev(thz_ctx* ctx, const float* d_cube, int n, int64_t P, float* avg /* [n] */);
/* The pixel means `ifft` takes (src/math_tools.rs:421-440) from the raw device cube: spectra are formed one chunk
 * of traces at a time and reduced, the [P][f] cubes never exist.  Any output may be NULL. */
int thz_mean_spectra_dev(thz_ctx* ctx, const float* d_raw, int64_t P, float* avg_fft /* [2 f] */,
                         float* avg_amp /* [f] */, float* avg_phase /* [f] */);

/* ---------------------------------------------------------------- deconvolution -------- */
/* PSF model as loaded from psf.npz (`load_psf`, src/io.rs:190-267; src/filters/psf.rs:7-22,
 * 202-207): natural cubic splines per segment a + b dx + c dx^2 + d dx^3 and the hybrid fit
 * w(f) = a/f + b + spline(f).  Arrays are borrowed. */
typedef struct {
  int n;                 /* knots */
  const float* knots;    /* [n]  THz */
  const float* values;   /* [n]  mm  */
  const float* coeff_a;  /* [n-1] (arrays of length n are accepted) */
  const float* coeff_b;
  const float* coeff_c;
  const float* coeff_d;
} thz_spline;
typedef struct { float base_a, base_b; thz_spline correction; } thz_hybrid_fit;
typedef struct { thz_hybrid_fit wx_fit, wy_fit; thz_spline x0_spline, y0_spline; } thz_psf;

#define THZ_MAX_PSF 255   /* PSF extent per axis, pixels */

/* `Deconvolution` parameters (src/filters/deconvolution.rs:217-240, defaults :725-733) */
typedef struct {
  int n_iterations;   /* 500 */
  int n_filters;      /* 25  */
  float start_freq;   /* 0.1 THz */
  float end_freq;     /* 10  THz */
  float win_width;    /* 0.5 THz */
} thz_deconv_params;

/* Everything `Deconvolution::filter` derives per band before touching the cube
 * (src/filters/deconvolution.rs:819-971): FIR taps, PSF factors (psf[i][j] = psf_x[i]*psf_y[j],
 * axis 0 <-> x, src/filters/psf.rs:305-311), iteration count, and which branch `convolve2d`
 * takes for this PSF (:484: area <= 256 -> direct *correlation*, else FFT *convolution*). */
typedef struct {
  float center_freq, wx, wy, x0, y0;
  int kx, ky;                 /* PSF extent along x (axis 0) and y (axis 1) */
  int n_iter;
  int direct;                 /* 1: correlation branch, 0: convolution branch */
  float psf_x[THZ_MAX_PSF];
  float psf_y[THZ_MAX_PSF];
  float fir[THZ_FIR_TAPS];
} thz_band_plan;

#define THZ_SKIP_NO_DXDY 2       /* the reference returns its input unchanged in these cases */
#define THZ_SKIP_NO_PSF 3        /* (src/filters/deconvolution.rs:781-812, 873-885)         */
#define THZ_SKIP_TOO_SMALL 4
#define THZ_SKIP_PSF_TOO_LARGE 5
/* NOT a reference skip: a band's PSF extends over more than THZ_MAX_PSF pixels per axis (the reference has no
 * cap).  The C++ / Rust hosts surface it through last_error instead of silently passing the data through. */
#define THZ_SKIP_PSF_UNSUPPORTED 6

/* `create_filter_bank` (src/filters/deconvolution.rs:160-211): n_filters x 499 Kaiser FIR
 * taps (designed in f64, stored as f32) and the log-spaced centre frequencies. */
int thz_fir_bank(int n_filters, double start_freq, double end_freq, double win_width, float t0, float t1,
                 float* filters /*[n_filters][499]*/, float* center_freqs /*[n_filters]*/);
/* psf.rs evaluators, exposed for tests */
float thz_hybrid_eval(const thz_hybrid_fit* fit, float f);
float thz_spline_eval_const_extrap(const thz_spline* s, float f);
/* Host part of `Deconvolution::filter`: returns THZ_OK and fills bands[n_filters], or one of
 * THZ_SKIP_* when the reference would return its input unchanged.  has_dxdy = 0 when the scan
 * has no dx / dy metadata. */
int thz_deconv_plan_bands(const thz_psf* psf, const thz_deconv_params* params, const float* time, int n,
                          int img_rows, int img_cols, int has_dxdy, float dx, float dy, thz_band_plan* bands);

/* Band energies: E[b][p] = sum_t (h_b * x[p])[t]^2 with the "same" alignment of `convolve1d`
 * (src/filters/deconvolution.rs:266-317, 574-609, 963-966).  d_energy is [n_bands][P].
 * Trace lengths: any n in [2, 7943] (zero-padded transform of the next power of two >= n + 249; powers of two
 * from 512 on run the N-point split / circular forms) and n = 8192 (N-point forms only). */
int thz_deconv_energies_dev(thz_ctx* ctx, const float* d_cube, int64_t P, int n, const thz_band_plan* bands,
                            int n_bands, float* d_energy);
/* `richardson_lucy` (src/filters/deconvolution.rs:620-712) on one [rows][cols] image with a
 * separable PSF: reflect pad, n_iter x { c = u (*) psf; r = d / (c + 1e-12); u *= r (*) mirror },
 * crop; then clamp >= 0 and, when d_gain != NULL, gain = sqrt(u / d) (:975, 990-993).
 * `direct` selects the correlation / convolution orientation of `convolve2d` (:484).
 * abort_flag (nullable) has the layout of Rust's `AtomicBool` (one byte, `abort_flag.as_ptr()`),
 * polled between launches like the cancellable loops poll it; progress (nullable) is called with
 * progress_base + progress_span * done/n_iter. */
int thz_rl_separable_dev(thz_ctx* ctx, const float* d_image, int rows, int cols, const float* psf_x, int kx,
                         const float* psf_y, int ky, int direct, int n_iter, float* d_deconvolved,
                         float* d_gain, const volatile uint8_t* abort_flag, thz_progress_fn progress,
                         void* progress_user, float progress_base, float progress_span);
/* Same iteration for an arbitrary dense PSF psf[kx][ky] (odd extents), tiled 2-D filtering. */
int thz_rl_dense_dev(thz_ctx* ctx, const float* d_image, int rows, int cols, const float* psf, int kx, int ky,
                     int direct, int n_iter, float* d_deconvolved, float* d_gain,
                     const volatile uint8_t* abort_flag);
/* One "same" 2-D filtering with zero boundary, exposed for tests: out = correlate(in, psf)
 * (direct != 0, src/filters/deconvolution.rs:432-458) or convolve(in, psf) (direct == 0, :489-544). */
int thz_conv2d_separable_dev(thz_ctx* ctx, const float* d_in, int rows, int cols, const float* psf_x, int kx,
                             const float* psf_y, int ky, int direct, float* d_out);
int thz_conv2d_dense_dev(thz_ctx* ctx, const float* d_in, int rows, int cols, const float* psf, int kx, int ky,
                         int direct, float* d_out);
/* Gain application and band sum (src/filters/deconvolution.rs:996-1012, 1030-1031):
 * out[p] = sum_b gain[b][p] * (h_b * x[p]); img[p] = sum_t out^2.  d_gain is [n_bands][P]. */
int thz_deconv_apply_dev(thz_ctx* ctx, const float* d_cube, const float* d_gain, int64_t P, int n,
                         const thz_band_plan* bands, int n_bands, float* d_out, float* d_img);
/* The whole filter on a device-resident cube [rows][cols][n] (slot 7 -> slot 8).  d_out may
 * alias d_cube.  Returns THZ_ABORTED when abort_flag became non-zero. */
int thz_deconvolution_dev(thz_ctx* ctx, const float* d_cube, int rows, int cols, int n,
                          const thz_band_plan* bands, int n_bands, float* d_out, float* d_img,
                          const volatile uint8_t* abort_flag, thz_progress_fn progress, void* progress_user);
/* CUDA-event timings of the last thz_deconvolution_dev call on this context, the per-filter
 * wall time the reference shows beside the filter header (src/data_thread.rs:1107, 1169-1184):
 * ms4 = {band energies, Richardson-Lucy, gain application, number of RL iterations run}. */
int thz_deconv_stage_ms(const thz_ctx* ctx, float* ms4);
/* Same call, per cube kernel (CUDA event pairs around each launch, read back after the call's final
 * synchronisation; the launches stay asynchronous): ms4 = {band-energy spectra pass, band-energy edge pass, gain-application edge corrections,
 * gain-application main pass}. */
int thz_deconv_kernel_ms(const thz_ctx* ctx, float* ms4);
/* Event pairs around the cube kernels of every call between begin and end (same four slots as
 * thz_deconv_kernel_ms); end synchronises the context. */
int thz_kernel_timing_begin(thz_ctx* ctx);
int thz_kernel_timing_end(thz_ctx* ctx, float* ms4 /* nullable */);
/* Host-pointer drop-in for `Deconvolution::filter`. */
int thz_deconvolution_host(thz_ctx* ctx, const float* cube, int rows, int cols, int n,
                           const thz_band_plan* bands, int n_bands, float* out, float* img,
                           const volatile uint8_t* abort_flag, thz_progress_fn progress, void* progress_user);

/* The whole chain in one call, host memory to host memory: the multipliers of thz_plan_trace
 * (slots 2..7) and, when n_bands > 0, `Deconvolution::filter` on the result; the cube
 * stays on the device in between (copies overlap the cube passes).  This is what
 * thzhost::ChainDriver::run_fused and bench.py's end-to-end figure use.  out may alias cube. */
int thz_chain_host(thz_ctx* ctx, const float* cube, int rows, int cols, int n, const thz_band_plan* bands,
                   int n_bands, float* out, float* img, const volatile uint8_t* abort_flag,
                   thz_progress_fn progress, void* progress_user);

/* The same chain on a device-resident cube (no copies): slots 2..7 of the chain fused with the band-energy
 * pass of `Deconvolution::filter` in ONE kernel (the filtered pair is still on chip when its FIR band energies are
 * formed, src/data_thread.rs:1090-1190 followed by src/filters/deconvolution.rs:963-966), then Richardson-Lucy
 * and the gain application.  thz_plan_trace(n, ...) must have been called; d_out may alias d_in; d_img is [rows*cols].
 * thz_deconv_stage_ms then reports {trace pass + band energies, Richardson-Lucy, gain application, iterations}.
 * Returns THZ_ABORTED when abort_flag became non-zero; d_out is then undefined (it may hold the hand-off between the
 * phases), so an in-place call that can be aborted should keep its input elsewhere. */
int thz_chain_dev(thz_ctx* ctx, const float* d_in, int rows, int cols, int n, const thz_band_plan* bands,
                  int n_bands, float* d_out, float* d_img, const volatile uint8_t* abort_flag,
                  thz_progress_fn progress, void* progress_user);
/* First part of thz_chain_dev alone (what a rank of a multi-GPU run calls before thz_slab_rl): d_out = filtered
 * traces, d_img = their intensities, d_energy [n_bands][P] = band energies of the filtered traces.  Power-of-two
 * n >= 512 run the fused kernel; other lengths (and THZ_CHAIN_FUSE=off) run thz_trace_fused_dev followed by
 * thz_deconv_energies_dev.  Results of the two forms agree to f32 rounding. */
int thz_chain_energies_dev(thz_ctx* ctx, const float* d_in, float* d_out, float* d_img, int64_t P, int n,
                           const thz_band_plan* bands, int n_bands, float* d_energy);
/* thz_chain_dev in two calls for a host that runs Richardson-Lucy itself in between (one rank of a multi-GPU run:
 * thz_slab_rl on d_energy -> d_gain).  d_work [P][n] carries the hand-off between the two halves in a private
 * layout (the spectra of the filtered trace pairs when the plan allows it -- pass C then needs no forward
 * transform -- otherwise the filtered traces) and must not be touched in between; d_out may alias d_work. */
int thz_chain_begin_dev(thz_ctx* ctx, const float* d_in, float* d_work, float* d_img, int64_t P, int n,
                        const thz_band_plan* bands, int n_bands, float* d_energy);
int thz_chain_end_dev(thz_ctx* ctx, const float* d_work, const float* d_gain, int64_t P, int n,
                      const thz_band_plan* bands, int n_bands, float* d_out, float* d_img);
/* Per-kernel CUDA-event sums of the last timed call (thz_chain_dev, thz_deconvolution_dev, or the calls between
 * thz_kernel_timing_begin/_end): ms5 = {band-energy spectra pass or the fused trace + energy kernel, band-energy
 * edge pass, gain-application edge corrections, gain-application main pass, trace pass when it ran on its own}. */
int thz_chain_kernel_ms(const thz_ctx* ctx, float* ms5);

/* thz_chain_host in two calls, for a host that runs the middle part itself (one process per GPU with
 * thz_slab_rl in between): begin = H2D chunks, fused trace pass, band energies (returns the device pointers of the
 * energies [n_bands][rows*cols] and of the gain buffer of the same shape); end = gain application, D2H chunks,
 * intensity map.  With n_bands = 0 begin already downloads the filtered cube and end only fetches the map. */
int thz_chain_host_begin(thz_ctx* ctx, const float* cube, int rows, int cols, int n, const thz_band_plan* bands,
                         int n_bands, float* out, float** d_energy, float** d_gain);
int thz_chain_host_end(thz_ctx* ctx, int rows, int cols, int n, const thz_band_plan* bands, int n_bands, float* out,
                       float* img);

/* ------------------------------------------------ several GPUs: row slabs ------------------ */
/* The cube is cut into contiguous row slabs over axis 0 (x), the axis the reference parallelises over
 * (src/math_tools.rs:333-340); the three cube passes are slab-local.  `richardson_lucy`
 * (src/filters/deconvolution.rs:620-712) runs on the same slabs of the reflect-padded band image: each rank
 * iterates its own rows, and the kx/2 boundary rows of `u` (before the first filtering of an iteration, :686-690)
 * and of the relative blur (before the second, :700-706) are stored straight into the neighbours' halo rows
 * over NVLink by the filtering kernels themselves, followed by a version flag; only the CTAs that read halo rows
 * wait for it.  No gather of the band images and no all-reduce of the gains: everything stays slab-local.
 *
 * One thz_slab per rank.  Ranks may be processes (one per GPU: exchange the 64-byte handles of
 * thz_slab_export() with any host transport and call thz_slab_connect_ipc) or live in one process
 * (thz_slab_connect_local; see also thz_group_* below, which drives all GPUs from one calling thread).
 *
 * thz_slab_plan is COLLECTIVE: every rank calls it with the same arguments while all ranks are idle, and a host
 * barrier separates it from the first thz_slab_rl (the arena holding the halos and flags is zeroed).  *changed:
 * 0 = same plan as before, 1 = new plan in the old arena, 2 = new arena (export + connect again).
 * row_bounds[world + 1]: image rows of rank r are [row_bounds[r], row_bounds[r + 1]).  Every slab must be
 * thicker than three PSF half-heights (THZ_EINVAL otherwise: use fewer ranks for that image). */
#define THZ_IPC_HANDLE_BYTES 64
typedef struct thz_slab thz_slab;
int thz_slab_create(thz_ctx* ctx, int rank, int world, thz_slab** out);
void thz_slab_destroy(thz_slab* slab);
int thz_slab_plan(thz_slab* slab, const int* row_bounds, int cols, const thz_band_plan* bands, int n_bands,
                  int* changed);
int thz_slab_export(thz_slab* slab, void* handle /* THZ_IPC_HANDLE_BYTES */);
int thz_slab_connect_ipc(thz_slab* slab, const void* handles /* [world][THZ_IPC_HANDLE_BYTES], by rank */);
int thz_slab_connect_local(thz_slab* slab, thz_slab* up /* rank - 1 or NULL */, thz_slab* down /* rank + 1 or NULL */);
int thz_slab_set_stream(thz_slab* slab, void* cuda_stream /* NULL = the context's stream */);
/* All bands of the plan on this rank's rows: d_energy [n_bands][bstride] (band energies of the slab, row-major
 * [rows_r][cols]) -> d_gain [n_bands][bstride] = sqrt(max(u, 0) / d), optionally the clamped estimate itself.
 * Asynchronous on the slab's stream; the abort flag is not polled in this form (a rank that stopped launching
 * would leave its neighbours waiting): abort between the phases instead. */
int thz_slab_rl(thz_slab* slab, const float* d_energy, int64_t bstride, float* d_gain, float* d_deconvolved);
/* The same run for n ranks that share ONE device (tests; hosts with fewer GPUs than slabs): all kernels go to
 * one stream in dependency order. */
int thz_slab_rl_serial(thz_slab* const* ranks, int n, const float* const* d_energy, const int64_t* bstride,
                       float* const* d_gain, float* const* d_deconvolved);
/* Synchronises the slab's stream; THZ_ECUDA when a halo wait timed out (a neighbour died). */
int thz_slab_status(thz_slab* slab);

/* One calling thread, several GPUs (the reference drives the chain from its single data thread,
 * src/data_thread.rs:162-174, 1090-1228).  A group owns one context and one slab per listed device, enables peer
 * access between neighbours, and runs each phase with one short-lived launching thread per device.  Listing the
 * same device several times emulates the slabs on one GPU (tests). */
typedef struct thz_group thz_group;
int thz_group_create(const int* devices, int n_devices, thz_group** out);
void thz_group_destroy(thz_group* group);
int thz_group_size(const thz_group* group);
thz_ctx* thz_group_ctx(thz_group* group, int rank);
const char* thz_group_last_error(const thz_group* group);
/* row_bounds[size + 1] of the partition the group uses for an image of `rows` rows */
int thz_group_row_bounds(const thz_group* group, int rows, int* row_bounds);
/* Halo-exchanged Richardson-Lucy of host band images energy[n_bands][rows][cols] -> gain[n_bands][rows][cols] */
int thz_group_rl_host(thz_group* group, const float* energy, int rows, int cols, const thz_band_plan* bands,
                      int n_bands, float* gain);
/* thz_chain_host over all devices of the group: the multipliers of thz_plan_trace, then (n_bands > 0)
 * `Deconvolution::filter`; host cube [rows][cols][n] in, filtered / deconvolved cube and intensity map out
 * (out may alias cube).  Every device uploads, filters and downloads its own row slab; the abort flag is polled
 * between the phases. */
int thz_group_chain_host(thz_group* group, const float* cube, int rows, int cols, int n, const float* m_pre,
                         const float* band, const float* m_post, const thz_band_plan* bands, int n_bands, float* out,
                         float* img, const volatile uint8_t* abort_flag, thz_progress_fn progress, void* progress_user);

/* ------------------------------------------------------- trace pass, host pointers ----- */
/* Same operators on host arrays (the reference's `ScannedImageFilterData` lives in host
 * memory).  Copies are chunked and overlapped with compute on three streams. */
int thz_trace_fused_host(thz_ctx* ctx, const float* in, float* out, float* img, int64_t P);
int thz_trace_forward_host(thz_ctx* ctx, const float* in, float* windowed, float* fft, float* amp,
                           float* phase, int64_t P);
int thz_trace_inverse_host(thz_ctx* ctx, const float* fft, int use_band, int use_post, float* out,
                           float* img, int64_t P);
int thz_time_multiply_host(thz_ctx* ctx, const float* in, const float* mult, int n, float* out, int64_t P);
int thz_band_apply_host(thz_ctx* ctx, float* fft, float* amp, int64_t P);   /* in place */
int thz_spectral_means_host(thz_ctx* ctx, const float* fft, const float* amp, const float* phase, int64_t P,
                            float* avg_fft, float* avg_amp, float* avg_phase);
/* img[p] = sum_t data[p][t]^2 (src/data_thread.rs:1288-1307) of a host cube */
int thz_intensity_host(thz_ctx* ctx, const float* data, int n, float* img, int64_t P);

/* ------------------------------------------------ chain driver (C++ host layer) -------- */
/* C handle over thzhost::ChainDriver (csrc/host/thz_host.hpp), the C++ mirror of the
 * reference's chain assembly (src/main.rs:194-268) and driver loop
 * (src/data_thread.rs:1090-1316) with the five shipped filters registered through the
 * `Filter` / `FilterRegistry` mirror.  Used by the Python tests to drive the chain exactly as
 * `data_thread` does; slots follow `filter_data_pipeline` (0 = loaded scan, i + 1 = output of
 * chain stage i). */
typedef struct thz_chain thz_chain;
int thz_chain_create(thz_ctx* ctx, thz_chain** out);
void thz_chain_destroy(thz_chain* chain);
int thz_chain_length(const thz_chain* chain);
const char* thz_chain_stage_name(thz_chain* chain, int i);
int thz_chain_set_config(thz_chain* chain, float window_lo, float window_hi, int window_type, int scale_factor);
int thz_chain_set_psf(thz_chain* chain, const thz_psf* psf);               /* ConfigCommand::ApplyPSF */
int thz_chain_set_param(thz_chain* chain, const char* filter_name, const char* param, double value);
int thz_chain_get_param(thz_chain* chain, const char* filter_name, const char* param, double* value);
int thz_chain_set_active(thz_chain* chain, const char* filter_name, int active);
int thz_chain_open(thz_chain* chain, const float* time, int n, const float* data, int width, int height,
                   int has_dxdy, float dx, float dy);                      /* ConfigCommand::OpenFile */
int thz_chain_run(thz_chain* chain, int start_idx, int run_deconvolution); /* UpdateType::Filter(start_idx) */
int thz_chain_run_fused(thz_chain* chain, int run_deconvolution);          /* same chain, one fused kernel */
void thz_chain_abort(thz_chain* chain, int value);                         /* the GUI's abort button */
int thz_chain_slot(thz_chain* chain, int slot, const float** data, const float** fft, const float** amp,
                   const float** phase, const float** img, const float** avg_fft, const float** avg_amp,
                   const float** avg_phase, int* n, int* f);
int thz_chain_fused_result(thz_chain* chain, const float** data, const float** img);
double thz_chain_filter_ms(thz_chain* chain, const char* filter_name);     /* filter_computation_time */

#ifdef __cplusplus
}
#endif
#endif /* THZGPU_H */
