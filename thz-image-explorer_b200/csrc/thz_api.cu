// thz_api.cu -- the C ABI (include/thzgpu.h): context, plans, memory, host-pointer pipelines.
#include "thz_internal.h"

#include <cstring>
#include <cstdlib>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <mutex>

static std::string g_create_err;
static std::mutex g_mu;

namespace thz {

int set_err(thz_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  else {
    std::lock_guard<std::mutex> lk(g_mu);
    g_create_err = msg;
  }
  return code;
}

int cuda_fail(thz_ctx* c, cudaError_t e, const char* what) {
  std::string m = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  return set_err(c, e == cudaErrorMemoryAllocation ? THZ_ENOMEM : THZ_ECUDA, m);
}

cudaError_t ensure_dynamic_smem(thz_ctx* c, const void* kernel, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> have;
  std::lock_guard<std::mutex> lk(mu);
  size_t& cur = have[{c->device, kernel}];
  if (cur >= bytes) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) cur = bytes;
  return e;
}

int get_tables(thz_ctx* c, int n, const FftTables** out) {
  auto it = c->tables.find(n);
  if (it != c->tables.end()) {
    *out = &it->second;
    return THZ_OK;
  }
  std::vector<float2> tw;
  if (build_twiddles(n, tw) != THZ_OK) return set_err(c, THZ_EINVAL, "unsupported FFT size");
  FftTables tb;
  tb.n = n;
  THZ_CUDA(c, cudaMalloc((void**)&tb.d_tw, tw.size() * sizeof(float2)));
  THZ_CUDA(c, cudaMemcpyAsync(tb.d_tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
  THZ_CUDA(c, cudaStreamSynchronize(c->stream));
  auto ins = c->tables.emplace(n, tb);
  *out = &ins.first->second;
  return THZ_OK;
}

int ensure_scratch(thz_ctx* c, size_t bytes) {
  if (c->scratch_bytes >= bytes) return THZ_OK;
  if (c->d_scratch) cudaFree(c->d_scratch);
  c->d_scratch = nullptr;
  c->scratch_bytes = 0;
  THZ_CUDA(c, cudaMalloc((void**)&c->d_scratch, bytes));
  c->scratch_bytes = bytes;
  return THZ_OK;
}

int ws_get(thz_ctx* c, int slot, size_t bytes, void** out) {
  auto& e = c->ws[slot];
  if (e.second < bytes) {
    if (e.first) {
      cudaError_t er = cudaFree(e.first);   // synchronises: nothing still uses the old buffer
      if (er != cudaSuccess) return cuda_fail(c, er, "cudaFree(workspace)");
    }
    e = {nullptr, 0};
    void* p = nullptr;
    cudaError_t er = cudaMalloc(&p, bytes);
    if (er != cudaSuccess) return cuda_fail(c, er, "cudaMalloc(workspace)");
    e = {p, bytes};
  }
  *out = e.first;
  return THZ_OK;
}

static int upload_vec(thz_ctx* c, float** d, const float* h, size_t n) {
  if (*d) {
    cudaFree(*d);
    *d = nullptr;
  }
  THZ_CUDA(c, cudaMalloc((void**)d, n * sizeof(float)));
  THZ_CUDA(c, cudaMemcpyAsync(*d, h, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  return THZ_OK;
}

static int ensure_stage(thz_ctx* c, size_t bytes) {
  if (c->stage_bytes >= bytes) return THZ_OK;
  for (int i = 0; i < kHostStreams; ++i) {
    if (c->d_stage[i]) cudaFree(c->d_stage[i]);
    c->d_stage[i] = nullptr;
  }
  c->stage_bytes = 0;
  for (int i = 0; i < kHostStreams; ++i) THZ_CUDA(c, cudaMalloc(&c->d_stage[i], bytes));
  c->stage_bytes = bytes;
  return THZ_OK;
}

}  // namespace thz

using namespace thz;

#define CHECK_CTX(c)                \
  do {                              \
    if (!(c)) return THZ_EINVAL;    \
    cudaError_t e_ = cudaSetDevice((c)->device); \
    if (e_ != cudaSuccess) return cuda_fail((c), e_, "cudaSetDevice"); \
  } while (0)

extern "C" {

int thz_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int thz_ctx_create(int device, thz_ctx** out) {
  if (!out) return THZ_EINVAL;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return set_err(nullptr, THZ_ECUDA, "no CUDA device available (libthzgpu has no CPU fallback)");
  }
  if (device < 0 || device >= n) return set_err(nullptr, THZ_EINVAL, "device index out of range");
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
  if (prop.major != 10) {
    char buf[160];
    snprintf(buf, sizeof buf, "device %d is sm_%d%d; libthzgpu is built for sm_100a only", device, prop.major,
             prop.minor);
    return set_err(nullptr, THZ_ECUDA, buf);
  }
  thz_ctx* c = new thz_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->unstaged_fir = true;
  if (const char* f = getenv("THZ_FIR_STAGING")) c->unstaged_fir = (strcmp(f, "on") != 0);
  if (const char* f = getenv("THZ_APPLY_FORM")) c->force_split_apply = (strcmp(f, "split") == 0);
  if (const char* f = getenv("THZ_EDGE_MMA")) c->edge_mma = (strcmp(f, "off") != 0);
  if (const char* f = getenv("THZ_RL_BATCH")) c->rl_batch = (strcmp(f, "off") != 0);
  if (const char* f = getenv("THZ_CHAIN_FUSE")) c->chain_fuse = (strcmp(f, "off") != 0);
  if (const char* f = getenv("THZ_CHAIN_SPECTRAL")) c->chain_spectral = (strcmp(f, "off") != 0);
  if (const char* f = getenv("THZ_CHAIN_EVEN")) c->chain_even_transform = (strcmp(f, "transform") == 0);
  if (const char* f = getenv("THZ_CHAIN_CHUNK_BYTES")) {
    const long long v = atoll(f);
    if (v >= 4096) c->host_chunk_bytes = (size_t)v;
  }
  e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete c;
    return cuda_fail(nullptr, e, "cudaStreamCreate");
  }
  for (int i = 0; i < kHostStreams; ++i) {
    e = cudaStreamCreateWithFlags(&c->hstream[i], cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      delete c;
      return cuda_fail(nullptr, e, "cudaStreamCreate");
    }
  }
  *out = c;
  return THZ_OK;
}

void thz_ctx_destroy(thz_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (auto& kv : c->tables)
    if (kv.second.d_tw) cudaFree(kv.second.d_tw);
  if (c->plan.d_m_pre) cudaFree(c->plan.d_m_pre);
  if (c->plan.d_m_post) cudaFree(c->plan.d_m_post);
  if (c->plan.d_band) cudaFree(c->plan.d_band);
  if (c->plan.d_hq) cudaFree(c->plan.d_hq);
  if (c->plan.d_chirp) cudaFree(c->plan.d_chirp);
  if (c->plan.d_bhat) cudaFree(c->plan.d_bhat);
  if (c->plan.d_hn) cudaFree(c->plan.d_hn);
  if (c->plan.d_ref_amp) cudaFree(c->plan.d_ref_amp);
  if (c->plan.d_ref_phase) cudaFree(c->plan.d_ref_phase);
  if (c->d_scratch) cudaFree(c->d_scratch);
  for (auto& kv : c->ws)
    if (kv.second.first) cudaFree(kv.second.first);
  for (int i = 0; i < kHostStreams; ++i) {
    if (c->d_stage[i]) cudaFree(c->d_stage[i]);
    if (c->hstream[i]) cudaStreamDestroy(c->hstream[i]);
  }
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char* thz_last_error(const thz_ctx* c) {
  if (c) return c->err.c_str();
  return g_create_err.c_str();
}

int thz_ctx_device(const thz_ctx* c) { return c ? c->device : -1; }
int thz_ctx_sm_count(const thz_ctx* c) { return c ? c->sm_count : 0; }
void* thz_ctx_stream(const thz_ctx* c) { return c ? (void*)c->stream : nullptr; }
int64_t thz_launch_count(const thz_ctx* c) { return c ? c->launches : 0; }

int thz_sync(thz_ctx* c) {
  CHECK_CTX(c);
  THZ_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < kHostStreams; ++i) THZ_CUDA(c, cudaStreamSynchronize(c->hstream[i]));
  return THZ_OK;
}

int thz_dev_alloc(thz_ctx* c, size_t bytes, void** d_ptr) {
  CHECK_CTX(c);
  if (!d_ptr) return set_err(c, THZ_EINVAL, "null out pointer");
  *d_ptr = nullptr;
  if (bytes == 0) bytes = 16;
  THZ_CUDA(c, cudaMalloc(d_ptr, bytes));
  return THZ_OK;
}

int thz_dev_free(thz_ctx* c, void* d_ptr) {
  CHECK_CTX(c);
  if (d_ptr) THZ_CUDA(c, cudaFree(d_ptr));
  return THZ_OK;
}

int thz_dev_memset(thz_ctx* c, void* d_ptr, int value, size_t bytes) {
  CHECK_CTX(c);
  THZ_CUDA(c, cudaMemsetAsync(d_ptr, value, bytes, c->stream));
  return THZ_OK;
}

int thz_copy_h2d(thz_ctx* c, void* d_dst, const void* src, size_t bytes) {
  CHECK_CTX(c);
  if (bytes == 0) return THZ_OK;
  THZ_CUDA(c, cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  THZ_CUDA(c, cudaStreamSynchronize(c->stream));
  return THZ_OK;
}

int thz_copy_d2h(thz_ctx* c, void* dst, const void* d_src, size_t bytes) {
  CHECK_CTX(c);
  if (bytes == 0) return THZ_OK;
  THZ_CUDA(c, cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, c->stream));
  THZ_CUDA(c, cudaStreamSynchronize(c->stream));
  return THZ_OK;
}

int thz_host_alloc(size_t bytes, void** ptr) {
  if (!ptr) return THZ_EINVAL;
  *ptr = nullptr;
  cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 16, cudaHostAllocDefault);
  if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaHostAlloc");
  return THZ_OK;
}

int thz_host_free(void* ptr) {
  if (ptr) {
    cudaError_t e = cudaFreeHost(ptr);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaFreeHost");
  }
  return THZ_OK;
}

int thz_generate_cube(thz_ctx* c, float* d_cube, int width, int height, int n, int row0, int total_width,
                      uint64_t seed, float t0, float dt, float noise) {
  CHECK_CTX(c);
  if (!d_cube || width < 0 || height < 0 || n <= 0) return set_err(c, THZ_EINVAL, "bad cube shape");
  return launch_generate(c, c->stream, d_cube, width, height, n, row0, total_width, seed, t0, dt, noise);
}

int thz_plan_trace(thz_ctx* c, int n, const float* m_pre, const float* band, const float* m_post) {
  CHECK_CTX(c);
  const bool pow2 = supported_n(n);
  if (!pow2 && !blue_supported(n))
    return set_err(c, THZ_EINVAL, "n must be a power of two in [64, 8192] or any other length in [2, 8192]");
  int rc;
  // the reference's stage functions rebuild their window per call (math_tools.rs:356-371) and so do the shims that
  // mirror them: an identical plan is recognised here and costs nothing (no synchronisation, no upload)
  {
    TracePlan& q = c->plan;
    auto same = [](const std::vector<float>& have, const float* want, size_t len) {
      if (!want) return have.empty();
      return have.size() == len && memcmp(have.data(), want, len * sizeof(float)) == 0;
    };
    if (q.n == n && q.valid && same(q.h_pre, m_pre, (size_t)n) && same(q.h_band, band, (size_t)n / 2 + 1) &&
        same(q.h_post, m_post, (size_t)n))
      return THZ_OK;
    q.valid = false;
    q.h_pre.assign(m_pre ? m_pre : nullptr, m_pre ? m_pre + n : nullptr);
    q.h_band.assign(band ? band : nullptr, band ? band + n / 2 + 1 : nullptr);
    q.h_post.assign(m_post ? m_post : nullptr, m_post ? m_post + n : nullptr);
  }
  // make sure no kernel still reads the old vectors
  THZ_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < kHostStreams; ++i) THZ_CUDA(c, cudaStreamSynchronize(c->hstream[i]));
  TracePlan& p = c->plan;
  p.n = n;
  p.has_pre = m_pre != nullptr;
  p.has_post = m_post != nullptr;
  auto ends_only = [n](const float* m) {
    if (!m || n < 64 || (n & (n - 1)) != 0) return false;
    for (int i = n / 16; i < n - n / 16; ++i)
      if (m[i] != 1.0f) return false;
    return true;
  };
  p.pre_ends_only = ends_only(m_pre);
  p.post_ends_only = ends_only(m_post);
  p.post_mode = 0;
  if (m_post) {
    for (int i = 0; i < n; ++i)
      if (m_post[i] != 1.0f) p.post_mode = std::max(p.post_mode, (i < 4 || i >= n - 4) ? 1 : 2);
  }
  p.has_band = band != nullptr;
  if (m_pre && (rc = upload_vec(c, &p.d_m_pre, m_pre, n)) != THZ_OK) return rc;
  if (m_post && (rc = upload_vec(c, &p.d_m_post, m_post, n)) != THZ_OK) return rc;
  if (band && (rc = upload_vec(c, &p.d_band, band, n / 2 + 1)) != THZ_OK) return rc;
  if (pow2) {
    const FftTables* tb = nullptr;
    if ((rc = get_tables(c, n, &tb)) != THZ_OK) return rc;
    p.blue_m = 0;
    std::vector<float> hq;
    if (build_hq(n, band, hq) != THZ_OK) return set_err(c, THZ_EINVAL, "unsupported n");
    if ((rc = upload_vec(c, &p.d_hq, hq.data(), n)) != THZ_OK) return rc;
    THZ_CUDA(c, cudaStreamSynchronize(c->stream));   // hq is a local
    p.valid = true;
    return THZ_OK;
  }
  // arbitrary length: chirp-z tables
  std::vector<float2> chirp, bhat;
  std::vector<float> hn;
  int m = 0;
  if (build_bluestein_tables(n, band, chirp, bhat, hn, m) != THZ_OK) return set_err(c, THZ_EINVAL, "unsupported n");
  const FftTables* tb = nullptr;
  if ((rc = get_tables(c, m > 8192 ? m / 2 : m, &tb)) != THZ_OK) return rc;   // 16384: two 8192-point sub-spectra
  if ((rc = upload_vec(c, (float**)&p.d_chirp, (const float*)chirp.data(), 2 * (size_t)n)) != THZ_OK) return rc;
  if ((rc = upload_vec(c, (float**)&p.d_bhat, (const float*)bhat.data(), 2 * bhat.size())) != THZ_OK) return rc;
  if ((rc = upload_vec(c, &p.d_hn, hn.data(), n)) != THZ_OK) return rc;
  THZ_CUDA(c, cudaStreamSynchronize(c->stream));
  p.blue_m = m;
  p.valid = true;
  return THZ_OK;
}

int thz_trace_fused_dev(thz_ctx* c, const float* d_in, float* d_out, float* d_img, int64_t P) {
  CHECK_CTX(c);
  return launch_trace_fused(c, c->stream, d_in, d_out, d_img, P);
}

int thz_trace_forward_dev(thz_ctx* c, const float* d_in, float* d_windowed, float* d_fft, float* d_amp,
                          float* d_phase, int64_t P) {
  CHECK_CTX(c);
  return launch_trace_forward(c, c->stream, d_in, d_windowed, (float2*)d_fft, d_amp, d_phase, P);
}

int thz_band_apply_dev(thz_ctx* c, float* d_fft, float* d_amp, int64_t P) {
  CHECK_CTX(c);
  return launch_band_apply(c, c->stream, (float2*)d_fft, d_amp, P);
}

int thz_trace_inverse_dev(thz_ctx* c, const float* d_fft, int use_band, int use_post, float* d_out, float* d_img,
                          int64_t P) {
  CHECK_CTX(c);
  return launch_trace_inverse(c, c->stream, (const float2*)d_fft, use_band != 0, use_post != 0, d_out, d_img, P);
}

static int upload_mult(thz_ctx* c, cudaStream_t s, const float* mult, int n, float** d) {
  *d = nullptr;
  if (!mult) return THZ_OK;
  void* p = nullptr;
  int rc = ws_get(c, WS_MULT, (size_t)n * sizeof(float), &p);
  if (rc != THZ_OK) return rc;
  THZ_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < kHostStreams; ++i) THZ_CUDA(c, cudaStreamSynchronize(c->hstream[i]));
  THZ_CUDA(c, cudaMemcpyAsync(p, mult, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s));
  THZ_CUDA(c, cudaStreamSynchronize(s));
  *d = (float*)p;
  return THZ_OK;
}

int thz_time_multiply_dev(thz_ctx* c, const float* d_in, const float* mult, int n, float* d_out, int64_t P) {
  CHECK_CTX(c);
  if (!d_in || !d_out || n <= 0) return set_err(c, THZ_EINVAL, "bad argument");
  float* d_m = nullptr;
  int rc = upload_mult(c, c->stream, mult, n, &d_m);
  if (rc != THZ_OK) return rc;
  return launch_time_multiply(c, c->stream, d_in, d_m, n, d_out, P, nullptr);
}

int thz_spectral_means(thz_ctx* c, const float* d_fft, const float* d_amp, const float* d_phase, int64_t P,
                       float* avg_fft, float* avg_amp, float* avg_phase) {
  CHECK_CTX(c);
  if (c->plan.n == 0) return set_err(c, THZ_ESTATE, "thz_plan_trace has not been called");
  if (P <= 0) return set_err(c, THZ_EINVAL, "P must be positive");
  const int F = c->plan.n / 2 + 1;
  int nb = c->sm_count * 4;
  if ((int64_t)nb > P) nb = (int)P;
  int rc = ensure_scratch(c, (size_t)nb * 2 * F * sizeof(float));
  if (rc != THZ_OK) return rc;
  std::vector<float> part((size_t)nb * 2 * F);
  struct Job { const float* d; int cols; float* out; };
  Job jobs[3] = {{d_fft, 2 * F, avg_fft}, {d_amp, F, avg_amp}, {d_phase, F, avg_phase}};
  for (const Job& j : jobs) {
    if (!j.d || !j.out) continue;
    rc = launch_column_sums(c, c->stream, j.d, P, j.cols, c->d_scratch, nb);
    if (rc != THZ_OK) return rc;
    THZ_CUDA(c, cudaMemcpyAsync(part.data(), c->d_scratch, (size_t)nb * j.cols * sizeof(float),
                                cudaMemcpyDeviceToHost, c->stream));
    THZ_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int col = 0; col < j.cols; ++col) {
      double acc = 0.0;
      for (int b = 0; b < nb; ++b) acc += (double)part[(size_t)b * j.cols + col];
      j.out[col] = (float)(acc / (double)P);
    }
  }
  return THZ_OK;
}

// ---------------------------------------------------------------- reference-normalised spectra (config 2)
int thz_plan_reference(thz_ctx* c, const float* ref_amp, const float* ref_phase, int f) {
  CHECK_CTX(c);
  THZ_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < kHostStreams; ++i) THZ_CUDA(c, cudaStreamSynchronize(c->hstream[i]));
  TracePlan& p = c->plan;
  if (!ref_amp || !ref_phase) {   // clear
    if (p.d_ref_amp) cudaFree(p.d_ref_amp);
    if (p.d_ref_phase) cudaFree(p.d_ref_phase);
    p.d_ref_amp = p.d_ref_phase = nullptr;
    p.ref_f = 0;
    return THZ_OK;
  }
  if (f < 1) return set_err(c, THZ_EINVAL, "bad reference length");
  int rc = upload_vec(c, &p.d_ref_amp, ref_amp, (size_t)f);
  if (rc == THZ_OK) rc = upload_vec(c, &p.d_ref_phase, ref_phase, (size_t)f);
  if (rc != THZ_OK) return rc;
  THZ_CUDA(c, cudaStreamSynchronize(c->stream));
  p.ref_f = f;
  return THZ_OK;
}

int thz_trace_forward_normalised_dev(thz_ctx* c, const float* d_in, float* d_windowed, float* d_fft, float* d_ratio,
                                     float* d_dphase, int64_t P) {
  CHECK_CTX(c);
  return launch_trace_forward(c, c->stream, d_in, d_windowed, (float2*)d_fft, d_ratio, d_dphase, P, true);
}

int thz_spectral_slice_dev(thz_ctx* c, const float* d_array, int f, int bin, float* d_map, int64_t P) {
  CHECK_CTX(c);
  if (!d_array || !d_map || f < 1 || bin < 0 || bin >= f || P < 0) return set_err(c, THZ_EINVAL, "bad argument");
  return launch_spectral_slice(c, c->stream, d_array, f, bin, d_map, P);
}

// ---------------------------------------------------------------- GUI hand-off from device-resident cubes
int thz_pixel_handoff_dev(thz_ctx* c, const float* d_raw, const float* d_filtered, int64_t P, int64_t pixel,
                          float* raw, float* filtered, float* fft, float* amp, float* phase) {
  CHECK_CTX(c);
  if (c->plan.n == 0) return set_err(c, THZ_ESTATE, "thz_plan_trace has not been called");
  if (pixel < 0 || pixel >= P) return set_err(c, THZ_EINVAL, "selected pixel out of bounds");   // data_thread.rs:1344-1356
  const int n = c->plan.n, F = n / 2 + 1;
  if (raw) {
    if (!d_raw) return set_err(c, THZ_EINVAL, "null raw cube");
    THZ_CUDA(c, cudaMemcpyAsync(raw, d_raw + pixel * n, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  }
  if (filtered) {
    if (!d_filtered) return set_err(c, THZ_EINVAL, "null filtered cube");
    THZ_CUDA(c, cudaMemcpyAsync(filtered, d_filtered + pixel * n, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost,
                                c->stream));
  }
  if (fft || amp || phase) {
    if (!d_raw) return set_err(c, THZ_EINVAL, "null raw cube");
    // the spectrum the reference plots is the one of slot fft + 1 (data_thread.rs:1364-1379): the windowed trace's
    // r2c with fft and amplitudes scaled by the band-pass (band_pass_fd.rs:155-159), phases untouched.  One
    // forward transform of one trace with the chain's own plan.
    void* ws = nullptr;
    int rc = ws_get(c, WS_HANDOFF, (size_t)(4 * F) * sizeof(float), &ws);
    if (rc != THZ_OK) return rc;
    float* d_f = (float*)ws;
    float* d_a = d_f + 2 * F;
    float* d_p = d_a + F;
    rc = launch_trace_forward(c, c->stream, d_raw + pixel * n, nullptr, (float2*)d_f, d_a, d_p, 1);
    if (rc == THZ_OK && c->plan.has_band) rc = launch_band_apply(c, c->stream, (float2*)d_f, d_a, 1);
    if (rc != THZ_OK) return rc;
    if (fft) THZ_CUDA(c, cudaMemcpyAsync(fft, d_f, (size_t)2 * F * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (amp) THZ_CUDA(c, cudaMemcpyAsync(amp, d_a, (size_t)F * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (phase) THZ_CUDA(c, cudaMemcpyAsync(phase, d_p, (size_t)F * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  }
  THZ_CUDA(c, cudaStreamSynchronize(c->stream));
  return THZ_OK;
}

// mean over the P traces of a [P][cols] device array -> host vector (per-block partial sums, combined in f64)
static int column_means(thz_ctx* c, const float* d_x, int64_t P, int cols, std::vector<double>& acc) {
  int nb = c->sm_count * 4;
  if ((int64_t)nb > P) nb = (int)P;
  int rc = ensure_scratch(c, (size_t)nb * cols * sizeof(float));
  if (rc != THZ_OK) return rc;
  std::vector<float> part((size_t)nb * cols);
  rc = launch_column_sums(c, c->stream, d_x, P, cols, c->d_scratch, nb);
  if (rc != THZ_OK) return rc;
  THZ_CUDA(c, cudaMemcpyAsync(part.data(), c->d_scratch, part.size() * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  THZ_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int b = 0; b < nb; ++b)
    for (int col = 0; col < cols; ++col) acc[col] += (double)part[(size_t)b * cols + col];
  return THZ_OK;
}

int thz_mean_trace_dev(thz_ctx* c, const float* d_cube, int n, int64_t P, float* avg) {
  CHECK_CTX(c);
  if (!d_cube || !avg || n < 1 || P < 1) return set_err(c, THZ_EINVAL, "bad argument");
  std::vector<double> acc((size_t)n, 0.0);
  int rc = column_means(c, d_cube, P, n, acc);
  if (rc != THZ_OK) return rc;
  for (int i = 0; i < n; ++i) avg[i] = (float)(acc[i] / (double)P);
  return THZ_OK;
}

int thz_mean_spectra_dev(thz_ctx* c, const float* d_raw, int64_t P, float* avg_fft, float* avg_amp, float* avg_phase) {
  CHECK_CTX(c);
  if (c->plan.n == 0) return set_err(c, THZ_ESTATE, "thz_plan_trace has not been called");
  if (!d_raw || P < 1) return set_err(c, THZ_EINVAL, "bad argument");
  const int n = c->plan.n, F = n / 2 + 1;
  // spectra of a chunk of traces at a time (<= 64 MiB of scratch), never the spectral cubes
  int64_t ct = ((int64_t)64 << 20) / ((int64_t)4 * F * sizeof(float));
  ct = std::max<int64_t>(2, ct & ~(int64_t)1);
  ct = std::min(ct, (P + 1) & ~(int64_t)1);
  void* ws = nullptr;
  int rc = ws_get(c, WS_MEANS_SPEC, (size_t)ct * 4 * F * sizeof(float), &ws);
  if (rc != THZ_OK) return rc;
  float* d_f = (float*)ws;
  float* d_a = d_f + (size_t)ct * 2 * F;
  float* d_p = d_a + (size_t)ct * F;
  std::vector<double> sf((size_t)2 * F, 0.0), sa((size_t)F, 0.0), sp((size_t)F, 0.0);
  for (int64_t p = 0; p < P; p += ct) {
    const int64_t np = std::min(ct, P - p);
    rc = launch_trace_forward(c, c->stream, d_raw + p * n, nullptr, avg_fft ? (float2*)d_f : nullptr,
                              avg_amp ? d_a : nullptr, avg_phase ? d_p : nullptr, np);
    // the means `ifft` takes are those of the band-passed slot (math_tools.rs:421-440 on the output of band_pass_fd)
    if (rc == THZ_OK && c->plan.has_band && (avg_fft || avg_amp))
      rc = launch_band_apply(c, c->stream, avg_fft ? (float2*)d_f : nullptr, avg_amp ? d_a : nullptr, np);
    if (rc == THZ_OK && avg_fft) rc = column_means(c, d_f, np, 2 * F, sf);
    if (rc == THZ_OK && avg_amp) rc = column_means(c, d_a, np, F, sa);
    if (rc == THZ_OK && avg_phase) rc = column_means(c, d_p, np, F, sp);
    if (rc != THZ_OK) return rc;
  }
  if (avg_fft) for (int i = 0; i < 2 * F; ++i) avg_fft[i] = (float)(sf[i] / (double)P);
  if (avg_amp) for (int i = 0; i < F; ++i) avg_amp[i] = (float)(sa[i] / (double)P);
  if (avg_phase) for (int i = 0; i < F; ++i) avg_phase[i] = (float)(sp[i] / (double)P);
  return THZ_OK;
}

// ---------------------------------------------------------------- host-pointer pipelines
// The cube is processed in chunks of whole trace pairs; chunk i uses stream i % 3 and that
// stream's staging buffer, so that H2D of chunk i+1, the kernel of chunk i and D2H of chunk
// i-1 overlap.
static int64_t chunk_traces(int n, int floats_per_trace_total) {
  // ~96 MiB of staging per stream
  int64_t t = ((int64_t)96 << 20) / ((int64_t)floats_per_trace_total * 4);
  t &= ~(int64_t)1023;
  if (t < 1024) t = 1024;
  (void)n;
  return t;
}

int thz_trace_fused_host(thz_ctx* c, const float* in, float* out, float* img, int64_t P) {
  CHECK_CTX(c);
  if (c->plan.n == 0) return set_err(c, THZ_ESTATE, "thz_plan_trace has not been called");
  if (P == 0) return THZ_OK;
  if (!in || !out) return set_err(c, THZ_EINVAL, "null pointer");
  const int n = c->plan.n;
  const int64_t ct = chunk_traces(n, n + 1);
  int rc = ensure_stage(c, (size_t)ct * (n + 1) * sizeof(float));
  if (rc != THZ_OK) return rc;
  int i = 0;
  for (int64_t p = 0; p < P; p += ct, ++i) {
    const int64_t np = std::min(ct, P - p);
    const int k = i % kHostStreams;
    cudaStream_t s = c->hstream[k];
    float* d_buf = (float*)c->d_stage[k];
    float* d_img = d_buf + (size_t)ct * n;
    THZ_CUDA(c, cudaMemcpyAsync(d_buf, in + p * n, (size_t)np * n * sizeof(float), cudaMemcpyHostToDevice, s));
    rc = launch_trace_fused(c, s, d_buf, d_buf, img ? d_img : nullptr, np);
    if (rc != THZ_OK) return rc;
    THZ_CUDA(c, cudaMemcpyAsync(out + p * n, d_buf, (size_t)np * n * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (img) THZ_CUDA(c, cudaMemcpyAsync(img + p, d_img, (size_t)np * sizeof(float), cudaMemcpyDeviceToHost, s));
  }
  for (int k = 0; k < kHostStreams; ++k) THZ_CUDA(c, cudaStreamSynchronize(c->hstream[k]));
  return THZ_OK;
}

int thz_trace_forward_host(thz_ctx* c, const float* in, float* windowed, float* fft, float* amp, float* phase,
                           int64_t P) {
  CHECK_CTX(c);
  if (c->plan.n == 0) return set_err(c, THZ_ESTATE, "thz_plan_trace has not been called");
  if (P == 0) return THZ_OK;
  if (!in) return set_err(c, THZ_EINVAL, "null pointer");
  const int n = c->plan.n, F = n / 2 + 1;
  const int per = n + 4 * F;   // data + fft(2F) + amp + phase
  const int64_t ct = chunk_traces(n, per);
  int rc = ensure_stage(c, (size_t)ct * per * sizeof(float));
  if (rc != THZ_OK) return rc;
  int i = 0;
  for (int64_t p = 0; p < P; p += ct, ++i) {
    const int64_t np = std::min(ct, P - p);
    const int k = i % kHostStreams;
    cudaStream_t s = c->hstream[k];
    float* d_data = (float*)c->d_stage[k];
    float* d_fft = d_data + (size_t)ct * n;
    float* d_amp = d_fft + (size_t)ct * 2 * F;
    float* d_ph = d_amp + (size_t)ct * F;
    THZ_CUDA(c, cudaMemcpyAsync(d_data, in + p * n, (size_t)np * n * sizeof(float), cudaMemcpyHostToDevice, s));
    rc = launch_trace_forward(c, s, d_data, windowed ? d_data : nullptr, fft ? (float2*)d_fft : nullptr,
                              amp ? d_amp : nullptr, phase ? d_ph : nullptr, np);
    if (rc != THZ_OK) return rc;
    if (windowed)
      THZ_CUDA(c, cudaMemcpyAsync(windowed + p * n, d_data, (size_t)np * n * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (fft)
      THZ_CUDA(c, cudaMemcpyAsync(fft + p * 2 * F, d_fft, (size_t)np * 2 * F * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (amp) THZ_CUDA(c, cudaMemcpyAsync(amp + p * F, d_amp, (size_t)np * F * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (phase)
      THZ_CUDA(c, cudaMemcpyAsync(phase + p * F, d_ph, (size_t)np * F * sizeof(float), cudaMemcpyDeviceToHost, s));
  }
  for (int k = 0; k < kHostStreams; ++k) THZ_CUDA(c, cudaStreamSynchronize(c->hstream[k]));
  return THZ_OK;
}

int thz_trace_inverse_host(thz_ctx* c, const float* fft, int use_band, int use_post, float* out, float* img,
                           int64_t P) {
  CHECK_CTX(c);
  if (c->plan.n == 0) return set_err(c, THZ_ESTATE, "thz_plan_trace has not been called");
  if (P == 0) return THZ_OK;
  if (!fft || !out) return set_err(c, THZ_EINVAL, "null pointer");
  const int n = c->plan.n, F = n / 2 + 1;
  const int per = n + 2 * F + 1;
  const int64_t ct = chunk_traces(n, per);
  int rc = ensure_stage(c, (size_t)ct * per * sizeof(float));
  if (rc != THZ_OK) return rc;
  int i = 0;
  for (int64_t p = 0; p < P; p += ct, ++i) {
    const int64_t np = std::min(ct, P - p);
    const int k = i % kHostStreams;
    cudaStream_t s = c->hstream[k];
    float* d_fft = (float*)c->d_stage[k];
    float* d_out = d_fft + (size_t)ct * 2 * F;
    float* d_img = d_out + (size_t)ct * n;
    THZ_CUDA(c, cudaMemcpyAsync(d_fft, fft + p * 2 * F, (size_t)np * 2 * F * sizeof(float), cudaMemcpyHostToDevice, s));
    rc = launch_trace_inverse(c, s, (const float2*)d_fft, use_band != 0, use_post != 0, d_out, img ? d_img : nullptr, np);
    if (rc != THZ_OK) return rc;
    THZ_CUDA(c, cudaMemcpyAsync(out + p * n, d_out, (size_t)np * n * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (img) THZ_CUDA(c, cudaMemcpyAsync(img + p, d_img, (size_t)np * sizeof(float), cudaMemcpyDeviceToHost, s));
  }
  for (int k = 0; k < kHostStreams; ++k) THZ_CUDA(c, cudaStreamSynchronize(c->hstream[k]));
  return THZ_OK;
}

// elementwise stages on host arrays: chunked over the three pipeline streams
static int elementwise_host(thz_ctx* c, const float* in, float* out, float* img, int64_t P, int n, const float* d_mult) {
  const int64_t ct = chunk_traces(n, n + 1);
  int rc = ensure_stage(c, (size_t)ct * (n + 1) * sizeof(float));
  if (rc != THZ_OK) return rc;
  int i = 0;
  for (int64_t p = 0; p < P; p += ct, ++i) {
    const int64_t np = std::min(ct, P - p);
    const int k = i % kHostStreams;
    cudaStream_t s = c->hstream[k];
    float* d_buf = (float*)c->d_stage[k];
    float* d_img = d_buf + (size_t)ct * n;
    THZ_CUDA(c, cudaMemcpyAsync(d_buf, in + p * n, (size_t)np * n * sizeof(float), cudaMemcpyHostToDevice, s));
    rc = launch_time_multiply(c, s, d_buf, d_mult, n, out ? d_buf : nullptr, np, img ? d_img : nullptr);
    if (rc != THZ_OK) return rc;
    if (out) THZ_CUDA(c, cudaMemcpyAsync(out + p * n, d_buf, (size_t)np * n * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (img) THZ_CUDA(c, cudaMemcpyAsync(img + p, d_img, (size_t)np * sizeof(float), cudaMemcpyDeviceToHost, s));
  }
  for (int k = 0; k < kHostStreams; ++k) THZ_CUDA(c, cudaStreamSynchronize(c->hstream[k]));
  return THZ_OK;
}

int thz_time_multiply_host(thz_ctx* c, const float* in, const float* mult, int n, float* out, int64_t P) {
  CHECK_CTX(c);
  if (P == 0) return THZ_OK;
  if (!in || !out || n <= 0) return set_err(c, THZ_EINVAL, "bad argument");
  float* d_m = nullptr;
  int rc = upload_mult(c, c->stream, mult, n, &d_m);
  if (rc != THZ_OK) return rc;
  return elementwise_host(c, in, out, nullptr, P, n, d_m);
}

int thz_intensity_host(thz_ctx* c, const float* data, int n, float* img, int64_t P) {
  CHECK_CTX(c);
  if (P == 0) return THZ_OK;
  if (!data || !img || n <= 0) return set_err(c, THZ_EINVAL, "bad argument");
  return elementwise_host(c, data, nullptr, img, P, n, nullptr);
}

int thz_band_apply_host(thz_ctx* c, float* fft, float* amp, int64_t P) {
  CHECK_CTX(c);
  if (c->plan.n == 0) return set_err(c, THZ_ESTATE, "thz_plan_trace has not been called");
  if (P == 0) return THZ_OK;
  const int F = c->plan.n / 2 + 1;
  const int per = 3 * F;
  const int64_t ct = chunk_traces(c->plan.n, per);
  int rc = ensure_stage(c, (size_t)ct * per * sizeof(float));
  if (rc != THZ_OK) return rc;
  int i = 0;
  for (int64_t p = 0; p < P; p += ct, ++i) {
    const int64_t np = std::min(ct, P - p);
    const int k = i % kHostStreams;
    cudaStream_t s = c->hstream[k];
    float* d_fft = (float*)c->d_stage[k];
    float* d_amp = d_fft + (size_t)ct * 2 * F;
    if (fft) THZ_CUDA(c, cudaMemcpyAsync(d_fft, fft + p * 2 * F, (size_t)np * 2 * F * sizeof(float), cudaMemcpyHostToDevice, s));
    if (amp) THZ_CUDA(c, cudaMemcpyAsync(d_amp, amp + p * F, (size_t)np * F * sizeof(float), cudaMemcpyHostToDevice, s));
    rc = launch_band_apply(c, s, fft ? (float2*)d_fft : nullptr, amp ? d_amp : nullptr, np);
    if (rc != THZ_OK) return rc;
    if (fft) THZ_CUDA(c, cudaMemcpyAsync(fft + p * 2 * F, d_fft, (size_t)np * 2 * F * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (amp) THZ_CUDA(c, cudaMemcpyAsync(amp + p * F, d_amp, (size_t)np * F * sizeof(float), cudaMemcpyDeviceToHost, s));
  }
  for (int k = 0; k < kHostStreams; ++k) THZ_CUDA(c, cudaStreamSynchronize(c->hstream[k]));
  return THZ_OK;
}

int thz_spectral_means_host(thz_ctx* c, const float* fft, const float* amp, const float* phase, int64_t P,
                            float* avg_fft, float* avg_amp, float* avg_phase) {
  CHECK_CTX(c);
  if (c->plan.n == 0) return set_err(c, THZ_ESTATE, "thz_plan_trace has not been called");
  if (P <= 0) return set_err(c, THZ_EINVAL, "P must be positive");
  const int F = c->plan.n / 2 + 1;
  const int nb = c->sm_count * 2;
  struct Job { const float* h; int cols; float* out; };
  Job jobs[3] = {{fft, 2 * F, avg_fft}, {amp, F, avg_amp}, {phase, F, avg_phase}};
  int rc = ensure_scratch(c, (size_t)nb * 2 * F * sizeof(float));
  if (rc != THZ_OK) return rc;
  std::vector<float> part((size_t)nb * 2 * F);
  for (const Job& j : jobs) {
    if (!j.h || !j.out) continue;
    const int64_t ct = chunk_traces(c->plan.n, j.cols);
    rc = ensure_stage(c, (size_t)ct * j.cols * sizeof(float));
    if (rc != THZ_OK) return rc;
    std::vector<double> acc(j.cols, 0.0);
    for (int64_t p = 0; p < P; p += ct) {
      const int64_t np = std::min(ct, P - p);
      float* d_buf = (float*)c->d_stage[0];
      THZ_CUDA(c, cudaMemcpyAsync(d_buf, j.h + p * j.cols, (size_t)np * j.cols * sizeof(float), cudaMemcpyHostToDevice, c->stream));
      const int nbb = (int)std::min<int64_t>(nb, np);
      rc = launch_column_sums(c, c->stream, d_buf, np, j.cols, c->d_scratch, nbb);
      if (rc != THZ_OK) return rc;
      THZ_CUDA(c, cudaMemcpyAsync(part.data(), c->d_scratch, (size_t)nbb * j.cols * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
      THZ_CUDA(c, cudaStreamSynchronize(c->stream));
      for (int b = 0; b < nbb; ++b)
        for (int col = 0; col < j.cols; ++col) acc[col] += (double)part[(size_t)b * j.cols + col];
    }
    for (int col = 0; col < j.cols; ++col) j.out[col] = (float)(acc[col] / (double)P);
  }
  return THZ_OK;
}

int thz_scale_blocks_dev(thz_ctx* c, const float* d_in, int width, int height, int zlen, int scale, float* d_out) {
  CHECK_CTX(c);
  if (!d_in || !d_out || scale < 1 || width < 0 || height < 0 || zlen < 0) return set_err(c, THZ_EINVAL, "bad argument");
  return launch_scale_blocks(c, c->stream, d_in, width, height, zlen, scale, d_out);
}

int thz_scale_blocks_host(thz_ctx* c, const float* in, int width, int height, int zlen, int scale, float* out) {
  CHECK_CTX(c);
  if (!in || !out || scale < 1) return set_err(c, THZ_EINVAL, "bad argument");
  const size_t nin = (size_t)width * height * zlen, nout = (size_t)(width / scale) * (height / scale) * zlen;
  if (nout == 0) return THZ_OK;
  void *pi = nullptr, *po = nullptr;
  int rc = ws_get(c, WS_SCALE_IN, nin * sizeof(float), &pi);
  if (rc == THZ_OK) rc = ws_get(c, WS_SCALE_OUT, nout * sizeof(float), &po);
  if (rc != THZ_OK) return rc;
  THZ_CUDA(c, cudaMemcpyAsync(pi, in, nin * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  rc = launch_scale_blocks(c, c->stream, (const float*)pi, width, height, zlen, scale, (float*)po);
  if (rc != THZ_OK) return rc;
  THZ_CUDA(c, cudaMemcpyAsync(out, po, nout * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  THZ_CUDA(c, cudaStreamSynchronize(c->stream));
  return THZ_OK;
}

int thz_bias_subtract_dev(thz_ctx* c, const float* d_in, int n, float* d_out, float* d_img, int64_t P) {
  CHECK_CTX(c);
  if (!d_in || !d_out || n <= 0) return set_err(c, THZ_EINVAL, "bad argument");
  return launch_bias_subtract(c, c->stream, d_in, n, d_out, P, d_img);
}

// point_in_polygon with the reference's `usize` arithmetic (release build: wrapping), src/math_tools.rs:573-591
static bool point_in_polygon(uint64_t x, uint64_t y, const std::vector<std::pair<uint64_t, uint64_t>>& poly) {
  bool inside = false;
  size_t j = poly.size() - 1;
  for (size_t i = 0; i < poly.size(); ++i) {
    const uint64_t xi = poly[i].first, yi = poly[i].second, xj = poly[j].first, yj = poly[j].second;
    if ((yi > y) != (yj > y)) {
      const uint64_t lim = (xj - xi) * (y - yi) / (yj - yi) + xi;   // unsigned wrap-around, as in Rust release
      if (x < lim) inside = !inside;
    }
    j = i;
  }
  return inside;
}

int thz_roi_average_dev(thz_ctx* c, const float* d_data, int dim0, int dim1, int zlen, const int64_t* poly_x,
                        const int64_t* poly_y, int n_points, int scaling, float* out) {
  CHECK_CTX(c);
  if (!d_data || !poly_x || !poly_y || !out || n_points < 1 || dim0 < 1 || dim1 < 1 || zlen < 1 || scaling < 1)
    return set_err(c, THZ_EINVAL, "bad argument");
  std::vector<std::pair<uint64_t, uint64_t>> poly(n_points);
  for (int i = 0; i < n_points; ++i) poly[i] = {(uint64_t)poly_x[i] / (uint64_t)scaling, (uint64_t)poly_y[i] / (uint64_t)scaling};
  const uint64_t x_size = (uint64_t)dim1, y_size = (uint64_t)dim0;
  uint64_t x_min = UINT64_MAX, y_min = UINT64_MAX, x_max = 0, y_max = 0;
  for (auto& p : poly) {
    x_min = std::min(x_min, p.first); y_min = std::min(y_min, p.second);
    x_max = std::max(x_max, p.first); y_max = std::max(y_max, p.second);
  }
  x_min = std::min(x_min, x_size - 1); y_min = std::min(y_min, y_size - 1);
  x_max = std::min(x_max, x_size - 1); y_max = std::min(y_max, y_size - 1);
  std::vector<int64_t> pix;
  for (uint64_t y = y_min; y <= y_max; ++y)
    for (uint64_t x = x_min; x <= x_max; ++x)
      if (point_in_polygon(x, y, poly)) pix.push_back((int64_t)((y_size - y - 1) * x_size + x));   // data[[y_size-y-1, x, z]]
  void *dp = nullptr, *dout = nullptr;
  int rc = ws_get(c, WS_ROI_PIX, std::max<size_t>(pix.size(), 1) * sizeof(int64_t), &dp);
  if (rc == THZ_OK) rc = ws_get(c, WS_ROI_OUT, (size_t)zlen * sizeof(float), &dout);
  if (rc != THZ_OK) return rc;
  if (!pix.empty())
    THZ_CUDA(c, cudaMemcpyAsync(dp, pix.data(), pix.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
  rc = launch_roi_average(c, c->stream, d_data, (const int64_t*)dp, (int)pix.size(), zlen, (float*)dout);
  if (rc != THZ_OK) return rc;
  THZ_CUDA(c, cudaMemcpyAsync(out, dout, (size_t)zlen * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  THZ_CUDA(c, cudaStreamSynchronize(c->stream));
  return THZ_OK;
}

int thz_tilt_shift_host(thz_ctx* c, const float* in, const float* taper, const int* insert, int n, int n_ext,
                        float* out, int64_t P) {
  CHECK_CTX(c);
  if (P == 0) return THZ_OK;
  if (!in || !taper || !insert || !out || n < 1 || n_ext < n) return set_err(c, THZ_EINVAL, "bad argument");
  void *pi = nullptr, *po = nullptr, *px = nullptr, *pt = nullptr;
  int rc = ws_get(c, WS_TILT_IN, (size_t)P * n * sizeof(float), &pi);
  if (rc == THZ_OK) rc = ws_get(c, WS_TILT_OUT, (size_t)P * n_ext * sizeof(float), &po);
  if (rc == THZ_OK) rc = ws_get(c, WS_TILT_IDX, (size_t)P * sizeof(int), &px);
  if (rc == THZ_OK) rc = ws_get(c, WS_TILT_TAPER, (size_t)n * sizeof(float), &pt);
  if (rc != THZ_OK) return rc;
  THZ_CUDA(c, cudaMemcpyAsync(pi, in, (size_t)P * n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  THZ_CUDA(c, cudaMemcpyAsync(px, insert, (size_t)P * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  THZ_CUDA(c, cudaMemcpyAsync(pt, taper, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  rc = launch_tilt_shift(c, c->stream, (const float*)pi, (const float*)pt, (const int*)px, n, n_ext, P, (float*)po);
  if (rc != THZ_OK) return rc;
  THZ_CUDA(c, cudaMemcpyAsync(out, po, (size_t)P * n_ext * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  THZ_CUDA(c, cudaStreamSynchronize(c->stream));
  return THZ_OK;
}

int thz_reference_pulse(thz_ctx* c, const float* scan_time, int n, const float* ref_time, const float* ref_signal,
                        int m, int window_type, float window_lo, float window_hi, float* signal_out, float* amp_out,
                        float* phase_out) {
  CHECK_CTX(c);
  if (!scan_time || !ref_time || !ref_signal || !signal_out || n < 2 || m < 1)
    return set_err(c, THZ_EINVAL, "bad argument");
  std::vector<float> ref(ref_signal, ref_signal + m);
  // resize / align (src/data_thread.rs:405-481)
  if (n != m || fabsf(scan_time[0] - ref_time[0]) > 1e-9f) {
    std::vector<float> nr((size_t)n, 0.f);
    if (n > 1 && m > 1) {
      const float ref_dt = ref_time[1] - ref_time[0];
      const float time_offset = scan_time[0] - ref_time[0];
      const long index_offset = (long)roundf(time_offset / ref_dt);
      const long src_start = index_offset > 0 ? index_offset : 0;
      const long dst_start = index_offset < 0 ? -index_offset : 0;
      const long copy_len = std::min(std::max((long)m - src_start, 0L), std::max((long)n - dst_start, 0L));
      for (long i = 0; i < copy_len; ++i) nr[(size_t)(dst_start + i)] = ref[(size_t)(src_start + i)];
    } else {
      for (int i = 0; i < std::min(n, m); ++i) nr[i] = ref[i];
    }
    ref.swap(nr);
  }
  // window on the reference file's own time axis; the zip stops at the shorter sequence (:489-511)
  std::vector<float> mult((size_t)m);
  int rc = thz_window_multiplier(window_type, ref_time, m, window_lo, window_hi, mult.data());
  if (rc != THZ_OK) return set_err(c, rc, "bad window type");
  for (int i = 0; i < std::min(n, m); ++i) ref[i] *= mult[i];
  memcpy(signal_out, ref.data(), (size_t)n * sizeof(float));
  if (!amp_out && !phase_out) return THZ_OK;
  // r2c + |s| + unwrap(arg s) with the scan's plan: the forward kernel on a 1 x 1 cube, no further window
  rc = thz_plan_trace(c, n, nullptr, nullptr, nullptr);
  if (rc != THZ_OK) return rc;
  return thz_trace_forward_host(c, ref.data(), nullptr, nullptr, amp_out, phase_out, 1);
}

int thz_voxel_opacity_dev(thz_ctx* c, const float* d_cube, int n, int64_t P, float opacity_threshold, float contrast,
                          float sigma, int radius, int64_t max_instances, float* d_opacity, float* effective_threshold) {
  CHECK_CTX(c);
  if (!d_cube || !d_opacity || n < 1 || radius < 0 || radius > 1024) return set_err(c, THZ_EINVAL, "bad argument");
  // gaussian_kernel1d (src/gui/threed_plot.rs:82-101), f32
  const int size = 2 * radius + 1;
  std::vector<float> kernel((size_t)size);
  const float sigma2 = 2.0f * sigma * sigma;
  float sum = 0.0f;
  for (int i = 0; i < size; ++i) {
    const float x = (float)i - (float)radius;
    kernel[i] = expf(-x * x / sigma2);
    sum += kernel[i];
  }
  for (float& v : kernel) v /= sum;
  void *pk = nullptr, *ph = nullptr;
  int rc = ws_get(c, WS_VOX_KERNEL, (size_t)size * sizeof(float), &pk);
  if (rc == THZ_OK) rc = ws_get(c, WS_VOX_HIST, 65536 * sizeof(unsigned long long), &ph);
  if (rc != THZ_OK) return rc;
  THZ_CUDA(c, cudaStreamSynchronize(c->stream));
  THZ_CUDA(c, cudaMemcpyAsync(pk, kernel.data(), (size_t)size * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  rc = launch_voxel_envelope(c, c->stream, d_cube, n, P, (const float*)pk, radius, contrast, opacity_threshold, d_opacity);
  if (rc != THZ_OK) return rc;
  THZ_CUDA(c, cudaStreamSynchronize(c->stream));
  if (!effective_threshold) return THZ_OK;
  // effective threshold = max_instances-th largest opacity when there are more voxels than that (:206-214);
  // opacities are in [0, 1]: their bit patterns order like unsigned integers -> two-pass radix select
  const int64_t total = P * n;
  *effective_threshold = 0.0f;
  if (total <= max_instances || max_instances < 1) return THZ_OK;
  std::vector<unsigned long long> hist(65536);
  unsigned prefix = 0;
  int64_t remaining = max_instances;   // rank from the top, 1-based
  for (int pass = 0; pass < 2; ++pass) {
    THZ_CUDA(c, cudaMemsetAsync(ph, 0, 65536 * sizeof(unsigned long long), c->stream));
    rc = launch_radix_hist(c, c->stream, d_opacity, total, pass, prefix, (unsigned long long*)ph);
    if (rc != THZ_OK) return rc;
    THZ_CUDA(c, cudaMemcpyAsync(hist.data(), ph, 65536 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    THZ_CUDA(c, cudaStreamSynchronize(c->stream));
    int b = 65535;
    for (; b >= 0; --b) {
      if ((int64_t)hist[b] >= remaining) break;
      remaining -= (int64_t)hist[b];
    }
    if (b < 0) b = 0;
    if (pass == 0) prefix = (unsigned)b;
    else {
      const unsigned bits = (prefix << 16) | (unsigned)b;
      memcpy(effective_threshold, &bits, sizeof(float));
    }
  }
  return THZ_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// FP32 issue-rate microbenchmark ("fp32 FMA peak to be measured by a microbenchmark in the same
// run"): the denominator of the Richardson-Lucy FLOP fraction and the evidence behind the choice between scalar
// and packed (f32x2) arithmetic in the transform kernels.  Every thread runs 16 independent dependency chains.
// mode 0: FFMA (three register operands)   1: fma.rn.f32x2 (packed, two lanes per instruction)
//      2: FADD                             3: add.rn.f32x2         4: FMUL      5: mul.rn.f32x2
// ---------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) k_fp32_rate(float* out, int iters, float a, float b) {
  constexpr bool PACKED = (MODE & 1) != 0;
  // per-thread (vector-register) operands, as in a butterfly: a uniform-register operand would time a cheaper form
  a += out[threadIdx.x] * 0.0f;
  b += out[threadIdx.x + 256] * 0.0f;
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = a + (float)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
      if constexpr (!PACKED) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if constexpr (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i]) : "f"(a), "f"(b));
          else if constexpr (MODE == 2) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(acc[i]) : "f"(b));
          else asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(acc[i]) : "f"(a));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          unsigned long long v, ca, cb;
          asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(acc[2 * i]), "f"(acc[2 * i + 1]));
          asm("mov.b64 %0, {%1, %1};" : "=l"(ca) : "f"(a));
          asm("mov.b64 %0, {%1, %1};" : "=l"(cb) : "f"(b));
          if constexpr (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(ca), "l"(cb));
          else if constexpr (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(cb));
          else asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(ca));
          asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * i]), "=f"(acc[2 * i + 1]) : "l"(v));
        }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += acc[i];
  if (s == 12345.678f) out[0] = s;   // never true: keeps the chains alive
}

extern "C" int thz_fp32_rate(thz_ctx* c, int mode, double* lane_ops_per_s) {
  CHECK_CTX(c);
  if (!lane_ops_per_s || mode < 0 || mode > 5) return set_err(c, THZ_EINVAL, "bad mode");
  void* scratch = nullptr;
  int rc = ws_get(c, WS_MULT, 4096, &scratch);
  if (rc != THZ_OK) return rc;
  const int iters = 4096, threads = 256, blocks = c->sm_count * 8;
  auto launch = [&](int it) {
    switch (mode) {
      case 0: k_fp32_rate<0><<<blocks, threads, 0, c->stream>>>((float*)scratch, it, 1.0000001f, 1e-9f); break;
      case 1: k_fp32_rate<1><<<blocks, threads, 0, c->stream>>>((float*)scratch, it, 1.0000001f, 1e-9f); break;
      case 2: k_fp32_rate<2><<<blocks, threads, 0, c->stream>>>((float*)scratch, it, 1.0000001f, 1e-9f); break;
      case 3: k_fp32_rate<3><<<blocks, threads, 0, c->stream>>>((float*)scratch, it, 1.0000001f, 1e-9f); break;
      case 4: k_fp32_rate<4><<<blocks, threads, 0, c->stream>>>((float*)scratch, it, 1.0000001f, 1e-9f); break;
      default: k_fp32_rate<5><<<blocks, threads, 0, c->stream>>>((float*)scratch, it, 1.0000001f, 1e-9f); break;
    }
    c->launches++;
  };
  launch(64);   // warm-up
  cudaEvent_t e0, e1;
  THZ_CUDA(c, cudaEventCreate(&e0));
  THZ_CUDA(c, cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0, c->stream);
    launch(iters);
    cudaEventRecord(e1, c->stream);
    THZ_CUDA(c, cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  // per thread and iteration: 4 reps x 16 instructions; a packed instruction works on two lanes
  const double lanes = (double)blocks * threads * (double)iters * 4.0 * 16.0 * ((mode & 1) ? 2.0 : 1.0);
  *lane_ops_per_s = lanes / ((double)best * 1e-3);
  return THZ_OK;
}
