// thz_group.cu -- several GPUs of one box behind ONE calling thread (the reference drives its whole
// chain from the single data thread, src/data_thread.rs:162-174, 1090).
//
// A group owns one context and one slab object per device.  The caller hands over the whole host cube; the
// group cuts it into row slabs over axis 0 (the axis the reference parallelises over, src/math_tools.rs:333-340)
// and runs, with one short-lived launching thread per device and phase,
//   in  : H2D chunks -> fused trace pass -> band energies           (slab-local, no communication)
//   RL  : halo-exchanged Richardson-Lucy on the slabs of the band images (peer stores over NVLink, thz_rl.cu)
//   out : gain application -> D2H chunks, intensity map             (slab-local)
// There is no gather of the band images and no reduction of the gains; the displayed map is written by every
// rank into its rows of the caller's `img`.
#include "thz_internal.h"

#include <algorithm>
#include <functional>
#include <string>
#include <thread>
#include <vector>

struct thz_group {
  std::vector<thz_ctx*> ctx;
  std::vector<thz_slab*> slab;
  std::string err;
  bool serial = false;   // several ranks on one device (tests): one stream, dependency order
};

namespace {

int group_fail(thz_group* g, int rank, int rc) {
  g->err = "rank " + std::to_string(rank) + ": " + (thz_last_error(g->ctx[rank]) ? thz_last_error(g->ctx[rank]) : "");
  return rc;
}

// fn(rank) on one thread per rank (the calling thread takes rank 0); first non-zero status wins
int parallel(thz_group* g, const std::function<int(int)>& fn) {
  const int n = (int)g->ctx.size();
  std::vector<int> rc(n, THZ_OK);
  if (g->serial || n == 1) {
    for (int r = 0; r < n; ++r) rc[r] = fn(r);
  } else {
    std::vector<std::thread> th;
    for (int r = 1; r < n; ++r) th.emplace_back([&, r] { rc[r] = fn(r); });
    rc[0] = fn(0);
    for (auto& t : th) t.join();
  }
  for (int r = 0; r < n; ++r)
    if (rc[r] != THZ_OK) return group_fail(g, r, rc[r]);
  return THZ_OK;
}

void bounds_of(int rows, int world, std::vector<int>& b) {
  b.assign(world + 1, 0);
  const int base = rows / world, rem = rows % world;
  for (int r = 0; r < world; ++r) b[r + 1] = b[r] + base + (r < rem ? 1 : 0);
}

// collective plan + local connection of all slabs; every rank is idle here
int plan_slabs(thz_group* g, const std::vector<int>& bounds, int cols, const thz_band_plan* bands, int n_bands) {
  const int n = (int)g->ctx.size();
  bool reconnect = false;
  for (int r = 0; r < n; ++r) {
    int rc = thz_sync(g->ctx[r]);
    if (rc != THZ_OK) return group_fail(g, r, rc);
  }
  for (int r = 0; r < n; ++r) {
    int changed = 0;
    int rc = thz_slab_plan(g->slab[r], bounds.data(), cols, bands, n_bands, &changed);
    if (rc != THZ_OK) return group_fail(g, r, rc);
    reconnect |= changed == 2;
  }
  if (reconnect)
    for (int r = 0; r < n; ++r) {
      int rc = thz_slab_connect_local(g->slab[r], r > 0 ? g->slab[r - 1] : nullptr, r + 1 < n ? g->slab[r + 1] : nullptr);
      if (rc != THZ_OK) return group_fail(g, r, rc);
    }
  return THZ_OK;
}

int run_rl(thz_group* g, const std::vector<const float*>& e, const std::vector<int64_t>& stride,
           const std::vector<float*>& gain) {
  const int n = (int)g->ctx.size();
  if (g->serial) {
    std::vector<float*> none(n, nullptr);
    int rc = thz_slab_rl_serial(g->slab.data(), n, e.data(), stride.data(), gain.data(), none.data());
    if (rc != THZ_OK) return group_fail(g, 0, rc);
    for (int r = 0; r < n; ++r) {
      rc = thz_slab_status(g->slab[r]);
      if (rc != THZ_OK) return group_fail(g, r, rc);
    }
    return THZ_OK;
  }
  return parallel(g, [&](int r) {
    int rc = thz_slab_rl(g->slab[r], e[r], stride[r], gain[r], nullptr);
    const int rs = thz_slab_status(g->slab[r]);   // always drain: the neighbours wait for this rank's rows
    return rc != THZ_OK ? rc : rs;
  });
}

}  // namespace

extern "C" {

int thz_group_create(const int* devices, int n_devices, thz_group** out) {
  if (!devices || n_devices < 1 || !out) return THZ_EINVAL;
  *out = nullptr;
  thz_group* g = new thz_group();
  for (int r = 0; r < n_devices; ++r)
    for (int q = 0; q < r; ++q) g->serial |= devices[q] == devices[r];
  if (g->serial)
    for (int r = 1; r < n_devices; ++r)
      if (devices[r] != devices[0]) {
        delete g;
        return thz::set_err(nullptr, THZ_EINVAL, "a device may be listed twice only when every rank uses the same one");
      }
  for (int r = 0; r < n_devices; ++r) {
    thz_ctx* c = nullptr;
    int rc = thz_ctx_create(devices[r], &c);
    if (rc == THZ_OK) {
      g->ctx.push_back(c);
      thz_slab* s = nullptr;
      rc = thz_slab_create(c, r, n_devices, &s);
      if (rc == THZ_OK) g->slab.push_back(s);
    }
    if (rc != THZ_OK) {
      thz_group_destroy(g);
      return rc;
    }
  }
  *out = g;
  return THZ_OK;
}

void thz_group_destroy(thz_group* g) {
  if (!g) return;
  for (thz_ctx* c : g->ctx) thz_sync(c);            // nobody pushes into an arena that is about to go
  for (thz_slab* s : g->slab) thz_slab_destroy(s);
  for (thz_ctx* c : g->ctx) thz_ctx_destroy(c);
  delete g;
}

int thz_group_size(const thz_group* g) { return g ? (int)g->ctx.size() : 0; }
thz_ctx* thz_group_ctx(thz_group* g, int rank) {
  return (g && rank >= 0 && rank < (int)g->ctx.size()) ? g->ctx[rank] : nullptr;
}
const char* thz_group_last_error(const thz_group* g) { return g ? g->err.c_str() : ""; }

int thz_group_row_bounds(const thz_group* g, int rows, int* bounds) {
  if (!g || !bounds || rows < 0) return THZ_EINVAL;
  std::vector<int> b;
  bounds_of(rows, (int)g->ctx.size(), b);
  std::copy(b.begin(), b.end(), bounds);
  return THZ_OK;
}

int thz_group_rl_host(thz_group* g, const float* energy, int rows, int cols, const thz_band_plan* bands, int n_bands,
                      float* gain) {
  if (!g || !energy || !gain || !bands || n_bands < 1) return THZ_EINVAL;
  const int n = (int)g->ctx.size();
  std::vector<int> bounds;
  bounds_of(rows, n, bounds);
  int rc = plan_slabs(g, bounds, cols, bands, n_bands);
  if (rc != THZ_OK) return rc;
  std::vector<const float*> e(n);
  std::vector<float*> gn(n);
  std::vector<int64_t> stride(n);
  for (int r = 0; r < n; ++r) {
    const int64_t pr = (int64_t)(bounds[r + 1] - bounds[r]) * cols;
    stride[r] = pr;
    void *pe = nullptr, *pg = nullptr;
    rc = thz_dev_alloc(g->ctx[r], (size_t)n_bands * pr * sizeof(float), &pe);
    if (rc == THZ_OK) rc = thz_dev_alloc(g->ctx[r], (size_t)n_bands * pr * sizeof(float), &pg);
    if (rc != THZ_OK) return group_fail(g, r, rc);
    e[r] = (const float*)pe;
    gn[r] = (float*)pg;
    for (int b = 0; b < n_bands && rc == THZ_OK; ++b)
      rc = thz_copy_h2d(g->ctx[r], (float*)pe + (size_t)b * pr,
                        energy + ((size_t)b * rows + bounds[r]) * cols, (size_t)pr * sizeof(float));
    if (rc != THZ_OK) return group_fail(g, r, rc);
  }
  rc = run_rl(g, e, stride, gn);
  for (int r = 0; r < n; ++r) {
    for (int b = 0; b < n_bands && rc == THZ_OK; ++b) {
      rc = thz_copy_d2h(g->ctx[r], gain + ((size_t)b * rows + bounds[r]) * cols, gn[r] + (size_t)b * stride[r],
                        (size_t)stride[r] * sizeof(float));
      if (rc != THZ_OK) group_fail(g, r, rc);
    }
    thz_dev_free(g->ctx[r], (void*)e[r]);
    thz_dev_free(g->ctx[r], gn[r]);
  }
  return rc;
}

int thz_group_chain_host(thz_group* g, const float* cube, int rows, int cols, int n, const float* m_pre,
                         const float* band, const float* m_post, const thz_band_plan* bands, int n_bands, float* out,
                         float* img, const volatile uint8_t* abort_flag, thz_progress_fn progress, void* progress_user) {
  if (!g || !cube || !out || rows < 1 || cols < 1) return THZ_EINVAL;
  const int world = (int)g->ctx.size();
  if (rows < world) {
    g->err = "fewer image rows than ranks";
    return THZ_EINVAL;
  }
  std::vector<int> bounds;
  bounds_of(rows, world, bounds);
  if (progress) progress(0.0f, progress_user);
  for (int r = 0; r < world; ++r) {
    int rc = thz_plan_trace(g->ctx[r], n, m_pre, band, m_post);
    if (rc != THZ_OK) return group_fail(g, r, rc);
  }
  if (n_bands > 0) {
    int rc = plan_slabs(g, bounds, cols, bands, n_bands);
    if (rc != THZ_OK) return rc;
  }
  std::vector<float*> d_energy(world, nullptr), d_gain(world, nullptr);
  const size_t row_floats = (size_t)cols * n;
  int rc = parallel(g, [&](int r) {
    thz_ctx* c = g->ctx[r];
    cudaSetDevice(c->device);
    const int64_t P = (int64_t)(bounds[r + 1] - bounds[r]) * cols;
    return thz::chain_pass_in(c, cube + (size_t)bounds[r] * row_floats, P, n, bands, n_bands,
                              out + (size_t)bounds[r] * row_floats, &d_energy[r], &d_gain[r]);
  });
  if (rc != THZ_OK) return rc;
  if (abort_flag && *abort_flag) return THZ_ABORTED;
  if (progress) progress(0.1f, progress_user);
  if (n_bands > 0) {
    std::vector<const float*> e(world);
    std::vector<int64_t> stride(world);
    for (int r = 0; r < world; ++r) {
      e[r] = d_energy[r];
      stride[r] = (int64_t)(bounds[r + 1] - bounds[r]) * cols;
    }
    rc = run_rl(g, e, stride, d_gain);
    if (rc != THZ_OK) return rc;
    if (abort_flag && *abort_flag) return THZ_ABORTED;
    if (progress) progress(0.9f, progress_user);
  }
  rc = parallel(g, [&](int r) {
    thz_ctx* c = g->ctx[r];
    cudaSetDevice(c->device);
    const int64_t P = (int64_t)(bounds[r + 1] - bounds[r]) * cols;
    return thz::chain_pass_out(c, P, n, bands, n_bands, out + (size_t)bounds[r] * row_floats,
                               img ? img + (size_t)bounds[r] * cols : nullptr);
  });
  if (rc == THZ_OK && progress) progress(1.0f, progress_user);
  return rc;
}

}  // extern "C"
