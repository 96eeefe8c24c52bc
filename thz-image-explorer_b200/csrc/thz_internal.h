// thz_internal.h -- context object and launcher prototypes shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <map>
#include <string>
#include <vector>

#include "../../include/thzgpu.h"

namespace thz {

// per-N device tables: twiddles for the DIF stages + digit-reversal helpers
struct FftTables {
  int n = 0;
  float2* d_tw = nullptr;       // TwTotal<N> entries
};

struct TracePlan {
  int n = 0;                    // samples per trace (power of two)
  float* d_m_pre = nullptr;     // [n] or null
  float* d_m_post = nullptr;    // [n] or null
  float* d_band = nullptr;      // [n/2+1] or null
  float* d_hq = nullptr;        // [n] band (or ones) / n in last-stage register order (fused kernel)
  bool has_pre = false, has_post = false, has_band = false;
  // the multiplier is exactly 1 on [n/16, n - n/16): only the first and the last register of a thread need it
  bool pre_ends_only = false, post_ends_only = false;
  // the gate after the inverse transform: 0 = none or all ones, 1 = differs from 1 only in the first / last four
  // samples (the default gate), 2 = general.  The fused trace + band-energy kernel derives the spectrum of the
  // gated trace from the filtered spectrum it already holds in modes 0 and 1 (thz_deconv.cu)
  int post_mode = 0;
  // traces whose length is not a power of two go through the chirp-z kernels (thz_bluestein.cu)
  int blue_m = 0;               // power-of-two transform size (>= 2n - 1), 0 = power-of-two plan
  float2* d_chirp = nullptr;    // [n]
  float2* d_bhat = nullptr;     // [blue_m]
  float* d_hn = nullptr;        // [n]
  // host copies of the planned vectors: thz_plan_trace returns at once when called again with the same plan
  bool valid = false;
  std::vector<float> h_pre, h_band, h_post;
  // reference pulse spectrum for the normalised forward outputs (thz_plan_reference)
  float* d_ref_amp = nullptr;   // [ref_f]
  float* d_ref_phase = nullptr; // [ref_f]
  int ref_f = 0;
};

constexpr int kHostStreams = 3;

// thz_deconv.cu (FIR passes A and C)
int deconv_energies(thz_ctx* c, cudaStream_t s, const float* d_cube, int64_t P, int n, const thz_band_plan* bands,
                    int B, float* d_energy, int64_t bstride = 0);
int deconv_apply(thz_ctx* c, cudaStream_t s, const float* d_cube, const float* d_gain, int64_t P, int n,
                 const thz_band_plan* bands, int B, float* d_out, float* d_img, int64_t bstride = 0, int lane = 0,
                 const float* d_edges = nullptr);
// d_edges != null selects the spectral hand-off: d_out receives the spectra of the filtered pairs (private layout),
// d_edges [P][512] their edge samples; deconv_apply must then be given the same d_edges
int chain_energies(thz_ctx* c, cudaStream_t s, const float* d_in, float* d_out, float* d_img, int64_t P, int n,
                   const thz_band_plan* bands, int B, float* d_energy, int64_t bstride = 0, float* d_edges = nullptr);
bool chain_spectral_ok(const thz_ctx* c, int n, int64_t P);
int chain_pass_in(thz_ctx* c, const float* cube, int64_t P, int n, const thz_band_plan* bands, int n_bands, float* out,
                  float** d_energy_out, float** d_gain_out);
int chain_pass_out(thz_ctx* c, int64_t P, int n, const thz_band_plan* bands, int n_bands, float* out, float* img);
// thz_rl.cu
int conv2d_once(thz_ctx* c, cudaStream_t s, const float* d_in, int rows, int cols, const float* psf_x, int kx,
                const float* psf_y, int ky, const float* dense, int direct, float* d_out);
int richardson_lucy(thz_ctx* c, cudaStream_t s, const float* d_image, int rows, int cols, const float* psf_x, int kx,
                    const float* psf_y, int ky, const float* dense, int direct, int n_iter, float* d_deconv,
                    float* d_gain, const volatile uint8_t* abort_flag, thz_progress_fn progress, void* puser,
                    float pbase, float pspan);
int richardson_lucy_bands(thz_ctx* c, cudaStream_t s, const float* d_energy, int64_t P, int rows, int cols,
                          const thz_band_plan* bands, int B, float* d_gain, const volatile uint8_t* abort_flag,
                          thz_progress_fn progress, void* puser, long* iterations_run);
// thz_edges_mma.cu
bool edges_mma_supported(int n);
int launch_fir_edges_mma(thz_ctx* c, cudaStream_t s, const float* d_cube, int64_t P, int n, const thz_band_plan* bands,
                         int B, float* d_energy, int64_t bstride, bool edge_rows = false);

}  // namespace thz

struct thz_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;                 // compute stream (device-pointer API)
  cudaStream_t hstream[thz::kHostStreams] = {};  // host-pointer API pipelines
  void* d_stage[thz::kHostStreams] = {};         // staging buffers of the host-pointer API
  size_t stage_bytes = 0;
  std::map<int, thz::FftTables> tables;
  std::map<const void*, int> occ;                // resident CTAs per SM, per kernel
  thz::TracePlan plan;
  std::string err;
  int64_t launches = 0;
  std::map<int, std::pair<void*, size_t>> ws;    // grow-only device workspace slots (freed at destroy)
  uint64_t fir_key = 0;                          // cache key of the uploaded FIR spectra (slot WS_FIR)
  int fir_m = 0;
  float stage_ms[4] = {0, 0, 0, 0};              // last thz_deconvolution_dev: energies, RL, apply, RL iterations
  float kernel_ms[5] = {0, 0, 0, 0, 0};          // same call, per kernel: energy spectra (or the fused trace + energy kernel), energy edges, apply edges, apply main, trace pass (when not fused)
  bool time_kernels = false;                     // set while thz_deconvolution_dev runs: event pairs around the cube kernels
  struct KernelEvent { int slot; cudaEvent_t e0, e1; };
  std::vector<KernelEvent> kernel_events;        // resolved after the call's final synchronisation
  bool force_split_apply = false;                // THZ_APPLY_FORM=split: zero-padded split form for pass C (A/B checks)
  bool edge_mma = true;                          // pass-A edge energies on the tensor cores for n >= 2048 (tcgen05, TF32); THZ_EDGE_MMA=off: transform kernel
  uint64_t edge_mma_key = 0;                     // taps the cached Toeplitz tiles (slot WS_EDGE_MMA) were built from
  int edge_mma_bands = 0;
  bool chain_fuse = true;                        // trace pass and band energies in one kernel (k_chain_energy_fused); THZ_CHAIN_FUSE=off: two passes
  bool chain_spectral = true;                    // whole-chain calls hand the spectra of the filtered pairs to pass C instead of the traces (THZ_CHAIN_SPECTRAL=off: traces)
  bool host_chain_spectral = false;              // mode chain_pass_in / thz_chain_begin_dev used, for the matching second half
  bool chain_even_transform = false;             // THZ_CHAIN_EVEN=transform: the fused kernel always transforms the stored trace for the even bins (A/B checks)
  bool rl_batch = true;                          // THZ_RL_BATCH=off: Richardson-Lucy band after band (A/B checks)
  bool unstaged_fir = true;                      // FIR passes read the cube directly so that L1 keeps the tables (THZ_FIR_STAGING=on: bulk-copy staging)
  size_t host_chunk_bytes = (size_t)256 << 20;   // chunk of the host-pointer pipelines (THZ_CHAIN_CHUNK_BYTES, tests shrink it)
  float* d_scratch = nullptr;                    // reductions
  size_t scratch_bytes = 0;
};

namespace thz {

int set_err(thz_ctx* c, int code, const std::string& msg);
int cuda_fail(thz_ctx* c, cudaError_t e, const char* what);
#define THZ_CUDA(ctx, call)                                              \
  do {                                                                   \
    cudaError_t e__ = (call);                                            \
    if (e__ != cudaSuccess) return ::thz::cuda_fail((ctx), e__, #call);  \
  } while (0)

// Opt a kernel in to at least `bytes` of dynamic shared memory.  The attribute belongs to the (device, kernel) pair,
// not to a context: the cache is process-wide and only ever raises the limit (a per-context cache let a second
// context lower it under the first one's feet).
cudaError_t ensure_dynamic_smem(thz_ctx* c, const void* kernel, size_t bytes);
int get_tables(thz_ctx* c, int n, const FftTables** out);
int ensure_scratch(thz_ctx* c, size_t bytes);
// grow-only workspace: returns a device buffer of at least `bytes` for `slot`
int ws_get(thz_ctx* c, int slot, size_t bytes, void** out);
enum { WS_FIR = 1, WS_ENERGY, WS_GAIN, WS_RL_D, WS_RL_U, WS_RL_R, WS_RL_TAPS, WS_CONV_A, WS_CONV_B, WS_HOST_CUBE, WS_HOST_IMG, WS_MULT, WS_SCALE_IN, WS_SCALE_OUT, WS_ROI_PIX, WS_ROI_OUT, WS_TILT_IN, WS_TILT_OUT, WS_TILT_IDX, WS_TILT_TAPER, WS_VOX_KERNEL, WS_VOX_HIST, WS_HANDOFF, WS_MEANS_SPEC, WS_RL_MULTI, WS_EDGE_MMA, WS_EDGE_ROWS,
       WS_EDGE_CORR /* + lane, lanes 0 .. kHostStreams */, WS_EDGE_CORR_LAST = WS_EDGE_CORR + kHostStreams, WS_END };

// thz_trace.cu
int launch_trace_fused(thz_ctx* c, cudaStream_t s, const float* d_in, float* d_out, float* d_img, int64_t P);
int launch_trace_forward(thz_ctx* c, cudaStream_t s, const float* d_in, float* d_win, float2* d_fft,
                         float* d_amp, float* d_phase, int64_t P, bool normalise = false);
int launch_spectral_slice(thz_ctx* c, cudaStream_t s, const float* d_a, int F, int bin, float* d_out, int64_t P);
int launch_trace_inverse(thz_ctx* c, cudaStream_t s, const float2* d_fft, bool use_band, bool use_post,
                         float* d_out, float* d_img, int64_t P);
int launch_time_multiply(thz_ctx* c, cudaStream_t s, const float* d_in, const float* d_mult, int n, float* d_out,
                         int64_t P, float* d_img);
int launch_scale_blocks(thz_ctx* c, cudaStream_t s, const float* d_in, int width, int height, int zlen, int sf,
                        float* d_out);
int launch_bias_subtract(thz_ctx* c, cudaStream_t s, const float* d_in, int n, float* d_out, int64_t P, float* d_img);
int launch_roi_average(thz_ctx* c, cudaStream_t s, const float* d_data, const int64_t* d_pix, int npix, int zlen,
                       float* d_out);
int launch_tilt_shift(thz_ctx* c, cudaStream_t s, const float* d_in, const float* d_taper, const int* d_insert, int n,
                      int n_ext, int64_t P, float* d_out);
int launch_voxel_envelope(thz_ctx* c, cudaStream_t s, const float* d_in, int n, int64_t P, const float* d_kernel,
                          int radius, float contrast, float thr, float* d_out);
int launch_radix_hist(thz_ctx* c, cudaStream_t s, const float* d_x, int64_t total, int pass, unsigned prefix,
                      unsigned long long* d_hist);
int launch_band_apply(thz_ctx* c, cudaStream_t s, float2* d_fft, float* d_amp, int64_t P);
int launch_column_sums(thz_ctx* c, cudaStream_t s, const float* d_x, int64_t rows, int cols, float* d_partials,
                       int nblocks);
int launch_generate(thz_ctx* c, cudaStream_t s, float* d_cube, int width, int height, int n, int row0,
                    int total_width, uint64_t seed, float t0, float dt, float noise);
int build_hq(int n, const float* band /*nullable, host*/, std::vector<float>& hq);
int build_twiddles(int n, std::vector<float2>& tw);
bool supported_n(int n);
// thz_bluestein.cu
bool blue_supported(int n);
int build_bluestein_tables(int n, const float* band, std::vector<float2>& chirp, std::vector<float2>& bhat,
                           std::vector<float>& hn, int& m_out);
int launch_blue_fused(thz_ctx* c, cudaStream_t s, const float* d_in, float* d_out, float* d_img, int64_t P);
int launch_blue_forward(thz_ctx* c, cudaStream_t s, const float* d_in, float* d_win, float2* d_fft, float* d_amp,
                        float* d_phase, int64_t P);
int launch_blue_inverse(thz_ctx* c, cudaStream_t s, const float2* d_fft, bool use_band, bool use_post, float* d_out,
                        float* d_img, int64_t P);
int richardson_lucy_bands(thz_ctx* c, cudaStream_t s, const float* d_energy, int64_t P, int rows, int cols,
                          const thz_band_plan* bands, int B, float* d_gain, const volatile uint8_t* abort_flag,
                          thz_progress_fn progress, void* puser, long* iterations_run);

}  // namespace thz
