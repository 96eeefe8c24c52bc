// thz_host.cpp -- C++ mirror of the reference's filter plugin API and chain driver, on top of
// the libthzgpu C ABI (see thz_host.hpp for the reference citations).
#include "thz_host.hpp"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <random>

namespace thzhost {

// ---------------------------------------------------------------------------------------
// PSF storage
// ---------------------------------------------------------------------------------------
static thz_spline spline_view(const std::vector<float> (&a)[6]) {
  thz_spline s;
  s.n = (int)a[0].size();
  s.knots = a[0].data();
  s.values = a[1].data();
  s.coeff_a = a[2].data();
  s.coeff_b = a[3].data();
  s.coeff_c = a[4].data();
  s.coeff_d = a[5].data();
  return s;
}
thz_psf PsfStorage::view() const {
  thz_psf p;
  p.wx_fit = {base_a[0], base_b[0], spline_view(arrays[0])};
  p.wy_fit = {base_a[1], base_b[1], spline_view(arrays[1])};
  p.x0_spline = spline_view(arrays[2]);
  p.y0_spline = spline_view(arrays[3]);
  return p;
}
static void spline_copy(std::vector<float> (&dst)[6], const thz_spline& s) {
  const float* src[6] = {s.knots, s.values, s.coeff_a, s.coeff_b, s.coeff_c, s.coeff_d};
  for (int k = 0; k < 6; ++k) dst[k].assign(src[k], src[k] + (k < 2 ? s.n : std::max(s.n - 1, 0)));
}
void PsfStorage::assign(const thz_psf& p) {
  spline_copy(arrays[0], p.wx_fit.correction);
  spline_copy(arrays[1], p.wy_fit.correction);
  spline_copy(arrays[2], p.x0_spline);
  spline_copy(arrays[3], p.y0_spline);
  base_a[0] = p.wx_fit.base_a; base_b[0] = p.wx_fit.base_b;
  base_a[1] = p.wy_fit.base_a; base_b[1] = p.wy_fit.base_b;
  loaded = p.wx_fit.correction.n > 0;
}

// ---------------------------------------------------------------------------------------
// registry
// ---------------------------------------------------------------------------------------
static std::string make_uuid() {
  static std::mt19937_64 rng{std::random_device{}()};
  char buf[40];
  const unsigned long long a = rng(), b = rng();
  snprintf(buf, sizeof buf, "%08llx-%04llx-4%03llx-%04llx-%012llx", a >> 32, (a >> 16) & 0xffff, a & 0xfff,
           (b >> 48) | 0x8000, b & 0xffffffffffffull);
  return buf;
}
FilterRegistry& FilterRegistry::global() {
  static FilterRegistry r;
  return r;
}
void FilterRegistry::add(std::unique_ptr<Filter> f) { filters.emplace(make_uuid(), std::move(f)); }
Filter* FilterRegistry::get_filter(const std::string& name) {
  for (auto& kv : filters)
    if (kv.second->config().name == name) return kv.second.get();
  return nullptr;
}

// ---------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------
static std::vector<float> frequency_of(const std::vector<float>& time) {
  std::vector<float> f(time.size() / 2 + 1);
  if (time.size() >= 2) thz_frequency_axis(time.data(), (int)time.size(), f.data());
  return f;
}

// power of two in [64, 8192] (fast kernels) or any length up to 4096 (chirp-z kernels)
static bool gpu_size_ok(size_t n) { return n >= 2 && n <= 8192; }

// a time-domain stage that multiplies every trace by one vector
static int multiply_stage(thz_ctx* ctx, const ScannedImageFilterData& in, const std::vector<float>& mult,
                          ScannedImageFilterData& out) {
  return thz_time_multiply_host(ctx, in.data.data(), mult.data(), (int)in.n(), out.data.data(), (int64_t)in.pixels());
}

// ---------------------------------------------------------------------------------------
// built-in stages (src/math_tools.rs:242-310, 330-398, 418-571)
// ---------------------------------------------------------------------------------------
ScannedImageFilterData scaling(thz_ctx* ctx, const ScannedImageFilterData& input, const ConfigContainer& config) {
  const int s = config.scale_factor;
  if (s <= 1) return input;                                   // math_tools.rs:244-246
  const size_t new_w = input.width / (size_t)s, new_h = input.height / (size_t)s;
  if (new_w == 0 || new_h == 0) return input;                  // "scaling is too large"
  ScannedImageFilterData out = input;
  out.width = new_w;
  out.height = new_h;
  if (out.dx) out.dx = *out.dx * (float)s;
  if (out.dy) out.dy = *out.dy * (float)s;
  const size_t n = input.n(), F = input.f(), P = new_w * new_h;
  const int W = (int)input.width, H = (int)input.height;
  out.data.assign(P * n, 0.f);
  thz_scale_blocks_host(ctx, input.data.data(), W, H, (int)n, s, out.data.data());
  if (input.amplitudes.size() == input.pixels() * F) {
    out.amplitudes.assign(P * F, 0.f);
    out.phases.assign(P * F, 0.f);
    out.fft.assign(P * F, {0.f, 0.f});
    thz_scale_blocks_host(ctx, input.amplitudes.data(), W, H, (int)F, s, out.amplitudes.data());
    thz_scale_blocks_host(ctx, input.phases.data(), W, H, (int)F, s, out.phases.data());
    thz_scale_blocks_host(ctx, reinterpret_cast<const float*>(input.fft.data()), W, H, (int)(2 * F), s,
                          reinterpret_cast<float*>(out.fft.data()));
  }
  out.img.assign(P, 0.f);   // the driver recomputes the intensity of the last slot
  return out;
}

ScannedImageFilterData fft(thz_ctx* ctx, const ScannedImageFilterData& input, const ConfigContainer& config) {
  ScannedImageFilterData out = input;   // the reference starts every stage with input.clone()
  if (!out.has_plan || !gpu_size_ok(out.n())) return out;
  const int n = (int)out.n();
  const size_t P = out.pixels(), F = out.f();
  std::vector<float> win(n);
  thz_window_multiplier((int)config.fft_window_type, out.time.data(), n, config.fft_window[0], config.fft_window[1],
                        win.data());
  if (thz_plan_trace(ctx, n, win.data(), nullptr, nullptr) != THZ_OK) return out;
  out.fft.assign(P * F, {0.f, 0.f});
  out.amplitudes.assign(P * F, 0.f);
  out.phases.assign(P * F, 0.f);
  // windowed data is left in `data` (math_tools.rs:356-371)
  thz_trace_forward_host(ctx, input.data.data(), out.data.data(), reinterpret_cast<float*>(out.fft.data()),
                         out.amplitudes.data(), out.phases.data(), (int64_t)P);
  return out;
}

ScannedImageFilterData ifft(thz_ctx* ctx, const ScannedImageFilterData& input, const ConfigContainer& config) {
  (void)config;
  ScannedImageFilterData out = input;
  const size_t P = out.pixels(), F = out.f();
  if (!gpu_size_ok(out.n()) || P == 0) return out;
  const int n = (int)out.n();
  if (thz_plan_trace(ctx, n, nullptr, nullptr, nullptr) != THZ_OK) return out;
  out.avg_fft.assign(F, {0.f, 0.f});
  out.avg_signal_fft.assign(F, 0.f);
  out.avg_phase_fft.assign(F, 0.f);
  thz_spectral_means_host(ctx, reinterpret_cast<const float*>(input.fft.data()), input.amplitudes.data(),
                          input.phases.data(), (int64_t)P, reinterpret_cast<float*>(out.avg_fft.data()),
                          out.avg_signal_fft.data(), out.avg_phase_fft.data());
  if (out.has_plan)
    thz_trace_inverse_host(ctx, reinterpret_cast<const float*>(input.fft.data()), 0, 0, out.data.data(), nullptr,
                           (int64_t)P);
  return out;
}

// ---------------------------------------------------------------------------------------
// the five shipped filters
// ---------------------------------------------------------------------------------------
#define THZ_FILTER_BOILERPLATE(T) \
  std::unique_ptr<Filter> clone_box() const override { return std::make_unique<T>(*this); }

// src/filters/tilt_compensation.rs
class TiltCompensation : public Filter {
 public:
  double tilt_x = 0.0, tilt_y = 0.0;
  void reset(const std::vector<float>&, const size_t[3]) override {}
  FilterConfig config() const override {
    return {"Tilt Compensation", "Compensate any misalignment of the scan along x and y.",
            FilterDomain::TimeBeforeFFTPrioFirst};
  }
  ScannedImageFilterData filter(const ScannedImageFilterData& input, GuiSettingsContainer&, ProgressLock&,
                                const std::atomic<bool>&) override {
    ScannedImageFilterData out = input;
    if (!(input.dx && input.dy) || input.time.empty()) return out;   // :111
    if (tilt_x != 0.0 || tilt_y != 0.0) {
      // per-pixel integer shift inside an axis extended by 2 * num_steps samples (:117-199)
      const int n = (int)input.n(), W = (int)input.width, H = (int)input.height;
      int num_steps = 0;
      thz_tilt_plan(input.time.data(), n, W, H, *input.dx, *input.dy, tilt_x, tilt_y, &num_steps, nullptr, nullptr);
      const int n_ext = n + 2 * num_steps;
      std::vector<float> time_ext((size_t)n_ext), taper((size_t)n);
      std::vector<int> insert(input.pixels());
      thz_tilt_plan(input.time.data(), n, W, H, *input.dx, *input.dy, tilt_x, tilt_y, &num_steps, time_ext.data(),
                    insert.data());
      thz_adapted_blackman(input.time.data(), n, 0.0f, 7.0f, taper.data());
      out.data.assign(input.pixels() * (size_t)n_ext, 0.f);
      thz_tilt_shift_host(env.ctx, input.data.data(), taper.data(), insert.data(), n, n_ext, out.data.data(),
                          (int64_t)input.pixels());
      out.time = time_ext;
      out.frequency = frequency_of(out.time);
      out.has_plan = true;
      return out;   // the driver zero-sizes the spectral cubes when the axis length changed (:1193-1227)
    }
    // 0 deg: no shift, no extension; every trace is tapered by the adapted Blackman (0, 7 ps) (:188)
    std::vector<float> taper(input.n());
    thz_adapted_blackman(input.time.data(), (int)input.n(), 0.0f, 7.0f, taper.data());
    multiply_stage(env.ctx, input, taper, out);
    out.frequency = frequency_of(out.time);   // re-planned (:204-216), same axis at 0 deg
    out.has_plan = true;
    return out;
  }
  bool set_param(const std::string& k, double v) override {
    if (k == "tilt_x") tilt_x = v; else if (k == "tilt_y") tilt_y = v; else return false;
    return true;
  }
  bool get_param(const std::string& k, double* v) const override {
    if (k == "tilt_x") *v = tilt_x; else if (k == "tilt_y") *v = tilt_y; else return false;
    return true;
  }
  THZ_FILTER_BOILERPLATE(TiltCompensation)
};
THZ_REGISTER_FILTER(TiltCompensation);

// src/filters/band_pass_td_before_fft.rs / band_pass_td_after_fft.rs (identical but for the
// default window width and the domain)
class TimeGateBase : public Filter {
 public:
  double low = 0.0, high = 0.0, window_width;
  explicit TimeGateBase(double ww) : window_width(ww) {}
  void reset(const std::vector<float>& time, const size_t[3]) override {   // :66-72
    low = time.empty() ? 0.0 : (double)time.front();
    high = time.empty() ? 0.0 : (double)time.back();
  }
  std::vector<float> multiplier(const std::vector<float>& time) {
    std::vector<float> m(time.size());
    thz_time_gate_multiplier(time.data(), (int)time.size(), &low, &high, window_width, m.data(), nullptr, nullptr);
    return m;
  }
  ScannedImageFilterData filter(const ScannedImageFilterData& input, GuiSettingsContainer&, ProgressLock& progress,
                                const std::atomic<bool>&) override {
    ScannedImageFilterData out = input;
    if (input.time.empty()) return out;
    multiply_stage(env.ctx, input, multiplier(input.time), out);
    progress(std::nullopt);
    return out;
  }
  bool set_param(const std::string& k, double v) override {
    if (k == "low") low = v; else if (k == "high") high = v; else if (k == "window_width") window_width = v; else return false;
    return true;
  }
  bool get_param(const std::string& k, double* v) const override {
    if (k == "low") *v = low; else if (k == "high") *v = high; else if (k == "window_width") *v = window_width; else return false;
    return true;
  }
};
class TimeDomainBandPassBeforeFFT : public TimeGateBase {
 public:
  TimeDomainBandPassBeforeFFT() : TimeGateBase(2.0) {}
  FilterConfig config() const override {
    return {"Time Domain Band Pass (before FFT)", "Band pass in the time domain, applied before the FFT.",
            FilterDomain::TimeBeforeFFT};
  }
  THZ_FILTER_BOILERPLATE(TimeDomainBandPassBeforeFFT)
};
THZ_REGISTER_FILTER(TimeDomainBandPassBeforeFFT);
class TimeDomainBandPassAfterFFT : public TimeGateBase {
 public:
  TimeDomainBandPassAfterFFT() : TimeGateBase(0.1) {}
  FilterConfig config() const override {
    return {"Time Domain Band Pass (after FFT)", "Band pass in the time domain, applied after the iFFT.",
            FilterDomain::TimeAfterFFT};
  }
  THZ_FILTER_BOILERPLATE(TimeDomainBandPassAfterFFT)
};
THZ_REGISTER_FILTER(TimeDomainBandPassAfterFFT);

// src/filters/band_pass_fd.rs
class FrequencyDomainBandPass : public Filter {
 public:
  double low = 0.2, high = 5.0, window_width = 0.1;   // :52-54
  void reset(const std::vector<float>&, const size_t[3]) override {}
  FilterConfig config() const override {
    return {"Frequency Domain Band Pass", "Band pass in the frequency domain.", FilterDomain::Frequency};
  }
  std::vector<float> multiplier(const std::vector<float>& freq) const {
    std::vector<float> m(freq.size());
    thz_band_pass_multiplier(freq.data(), (int)freq.size(), low, high, window_width, m.data(), nullptr, nullptr);
    return m;
  }
  ScannedImageFilterData filter(const ScannedImageFilterData& input, GuiSettingsContainer&, ProgressLock& progress,
                                const std::atomic<bool>&) override {
    ScannedImageFilterData out = input;
    if (input.frequency.empty() || input.fft.empty() || !gpu_size_ok(input.n())) return out;
    const std::vector<float> band = multiplier(input.frequency);
    if (thz_plan_trace(env.ctx, (int)input.n(), nullptr, band.data(), nullptr) == THZ_OK)
      thz_band_apply_host(env.ctx, reinterpret_cast<float*>(out.fft.data()), out.amplitudes.data(),
                          (int64_t)input.pixels());   // phases untouched (:155-159)
    progress(std::nullopt);
    return out;
  }
  bool set_param(const std::string& k, double v) override {
    if (k == "low") low = v; else if (k == "high") high = v; else if (k == "window_width") window_width = v; else return false;
    return true;
  }
  bool get_param(const std::string& k, double* v) const override {
    if (k == "low") *v = low; else if (k == "high") *v = high; else if (k == "window_width") *v = window_width; else return false;
    return true;
  }
  THZ_FILTER_BOILERPLATE(FrequencyDomainBandPass)
};
THZ_REGISTER_FILTER(FrequencyDomainBandPass);

// src/filters/deconvolution.rs
class Deconvolution : public Filter {
 public:
  int n_iterations = 500, n_filters = 25;              // :725-733
  float start_freq = 0.1f, end_freq = 10.0f, win_width = 0.5f;
  void reset(const std::vector<float>&, const size_t[3]) override {}
  FilterConfig config() const override {
    return {"Deconvolution", "Frequency dependent Richardson-Lucy deconvolution with the beam PSF.",
            FilterDomain::TimeAfterFFTPrioLast};
  }
  ScannedImageFilterData filter(const ScannedImageFilterData& input, GuiSettingsContainer& gui, ProgressLock& progress,
                                const std::atomic<bool>& abort_flag) override {
    progress(0.0f);
    ScannedImageFilterData out = input;
    const thz_psf psf = gui.psf.view();
    thz_deconv_params prm{n_iterations, n_filters, start_freq, end_freq, win_width};
    std::vector<thz_band_plan> bands((size_t)std::max(n_filters, 1));
    const bool has = input.dx && input.dy;
    const int rc = thz_deconv_plan_bands(gui.psf.loaded ? &psf : nullptr, &prm, input.time.data(), (int)input.n(),
                                         (int)input.width, (int)input.height, has ? 1 : 0, has ? *input.dx : 0.f,
                                         has ? *input.dy : 0.f, bands.data());
    if (rc != THZ_OK) {   // the reference logs and returns input.clone() (:781-812, 873-885)
      if (rc == THZ_SKIP_PSF_UNSUPPORTED || rc < 0)
        env.last_error = "Deconvolution skipped: PSF extent above THZ_MAX_PSF pixels or bad parameters (library limit, "
                         "the reference would have run)";
      progress(std::nullopt);
      return out;
    }
    struct Cb {
      ProgressLock* p;
      static void fn(float f, void* u) { (*static_cast<Cb*>(u)->p)(f); }
    } cb{&progress};
    // the AtomicBool is polled through the C ABI as a volatile byte
    const volatile uint8_t* abort_byte = reinterpret_cast<const volatile uint8_t*>(&abort_flag);
    out.img.assign(input.pixels(), 0.f);
    const int rcd = thz_deconvolution_host(env.ctx, input.data.data(), (int)input.width, (int)input.height,
                                           (int)input.n(), bands.data(), n_filters, out.data.data(), out.img.data(),
                                           abort_byte, &Cb::fn, &cb);
    if (rcd != THZ_OK) out = input;   // aborted or failed: keep the input (:1016-1024)
    progress(std::nullopt);
    return out;
  }
  bool set_param(const std::string& k, double v) override {
    if (k == "n_iterations") n_iterations = (int)v; else if (k == "n_filters") n_filters = (int)v;
    else if (k == "start_freq") start_freq = (float)v; else if (k == "end_freq") end_freq = (float)v;
    else if (k == "win_width") win_width = (float)v; else return false;
    return true;
  }
  bool get_param(const std::string& k, double* v) const override {
    if (k == "n_iterations") *v = n_iterations; else if (k == "n_filters") *v = n_filters;
    else if (k == "start_freq") *v = start_freq; else if (k == "end_freq") *v = end_freq;
    else if (k == "win_width") *v = win_width; else return false;
    return true;
  }
  THZ_FILTER_BOILERPLATE(Deconvolution)
};
THZ_REGISTER_FILTER(Deconvolution);

// ---------------------------------------------------------------------------------------
// chain assembly (src/main.rs:194-268) and driver (src/data_thread.rs:1023-1316)
// ---------------------------------------------------------------------------------------
ChainDriver::ChainDriver(thz_ctx* ctx) : ctx_(ctx) {
  auto& reg = FilterRegistry::global();
  std::vector<std::string> ordered;
  scaling_index = ordered.size();
  ordered.push_back("scaling");
  for (FilterDomain d : {FilterDomain::TimeBeforeFFTPrioFirst, FilterDomain::TimeBeforeFFT})
    for (auto& kv : reg.filters)
      if (kv.second->config().domain == d) ordered.push_back(kv.first);
  fft_index = ordered.size();
  ordered.push_back("fft");
  for (auto& kv : reg.filters)
    if (kv.second->config().domain == FilterDomain::Frequency) ordered.push_back(kv.first);
  ifft_index = ordered.size();
  ordered.push_back("ifft");
  for (FilterDomain d : {FilterDomain::TimeAfterFFT, FilterDomain::TimeAfterFFTPrioLast})
    for (auto& kv : reg.filters)
      if (kv.second->config().domain == d) ordered.push_back(kv.first);
  for (size_t i = 0; i < ordered.size(); ++i) {
    filter_chain.push_back(ordered[i]);
    filter_uuid_to_index[ordered[i]] = i + 1;
  }
  for (auto& kv : reg.filters) {
    // deconvolution filters are disabled by default (main.rs:254-258)
    filters_active[kv.first] = kv.second->config().name.find("Deconvolution") == std::string::npos;
    filter_computation_time[kv.first] = std::chrono::duration<double>(0);
    filters_[kv.first] = kv.second->clone_box();
    filters_[kv.first]->env.ctx = ctx_;
  }
  filter_data_pipeline.resize(ordered.size() + 1);
}

std::string ChainDriver::uuid_of(const std::string& name) const {
  for (auto& kv : filters_)
    if (kv.second->config().name == name) return kv.first;
  return "";
}
Filter* ChainDriver::filter_by_name(const std::string& name) {
  const std::string u = uuid_of(name);
  return u.empty() ? nullptr : filters_[u].get();
}

void ChainDriver::open(const std::vector<float>& time, const float* data, size_t width, size_t height,
                       std::optional<float> dx, std::optional<float> dy) {
  ScannedImageFilterData s;
  s.time = time;
  s.frequency = frequency_of(time);
  s.width = width;
  s.height = height;
  s.dx = dx;
  s.dy = dy;
  s.has_plan = time.size() >= 2;
  const size_t P = width * height, n = time.size(), F = s.frequency.size();
  s.data.assign(data, data + P * n);
  s.fft.assign(P * F, {0.f, 0.f});
  s.amplitudes.assign(P * F, 0.f);
  s.phases.assign(P * F, 0.f);
  s.img.assign(P, 0.f);
  if (P) thz_intensity_host(ctx_, s.data.data(), (int)n, s.img.data(), (int64_t)P);
  filter_data_pipeline[0] = std::move(s);
  const size_t shape[3] = {width, height, n};
  for (auto& kv : filters_) kv.second->reset(time, shape);   // data_thread.rs:1027-1060
}

int ChainDriver::run(size_t start_idx, bool run_deconvolution) {
  if (start_idx < 1) start_idx = 1;
  auto& fd = filter_data_pipeline;
  for (size_t i = start_idx - 1; i < filter_chain.size(); ++i) {
    const std::string& id = filter_chain[i];
    const size_t out_idx = filter_uuid_to_index[id];
    const size_t in_idx = (i == 0) ? 0 : filter_uuid_to_index[filter_chain[i - 1]];
    if (fd[in_idx].time.empty()) continue;   // "Input data for filter is empty, skipping"
    const auto t0 = std::chrono::steady_clock::now();
    if (id == "scaling") fd[out_idx] = scaling(ctx_, fd[in_idx], config);
    else if (id == "fft") fd[out_idx] = fft(ctx_, fd[in_idx], config);
    else if (id == "ifft") fd[out_idx] = ifft(ctx_, fd[in_idx], config);
    else {
      Filter* f = filters_[id].get();
      const bool active = filters_active[id];
      const bool deconv = f->config().name.find("Deconvolution") != std::string::npos;
      if (!deconv) run_deconvolution = false;   // data_thread.rs:1139-1150 (cleared by any earlier filter)
      if (active && !(deconv && !run_deconvolution)) {
        ProgressLock progress = [this](std::optional<float> p) { last_progress = p; };
        f->env.last_error.clear();
        fd[out_idx] = f->filter(fd[in_idx], gui_settings, progress, abort_flag);
        if (!f->env.last_error.empty()) last_error = f->env.last_error;
        f->show_data(fd[out_idx]);
        filter_computation_time[id] = std::chrono::steady_clock::now() - t0;
      } else {
        fd[out_idx] = fd[in_idx];
      }
    }
    if (fd[in_idx].n() != fd[out_idx].n()) {   // re-plan when a stage changed the time axis (:1193-1227)
      auto& o = fd[out_idx];
      o.frequency = frequency_of(o.time);
      o.has_plan = true;
      const size_t P = o.pixels(), F = o.f();
      o.fft.assign(P * F, {0.f, 0.f});
      o.amplitudes.assign(P * F, 0.f);
      o.phases.assign(P * F, 0.f);
    }
  }
  // intensity image of the last slot (data_thread.rs:1243-1308)
  auto& last = fd[filter_uuid_to_index[filter_chain.back()]];
  if (last.pixels()) {
    last.img.assign(last.pixels(), 0.f);
    thz_intensity_host(ctx_, last.data.data(), (int)last.n(), last.img.data(), (int64_t)last.pixels());
  }
  return THZ_OK;
}

int ChainDriver::run_fused(bool run_deconvolution) {
  // the first chain slot is `scaling` (math_tools.rs:242-310): with scale_factor > 1 the fused kernels run on the
  // block-mean cube, with its dimensions and dx * s, dy * s, exactly as the staged run does
  ScannedImageFilterData scaled;
  const bool do_scale = config.scale_factor > 1 && filter_data_pipeline[0].width / (size_t)config.scale_factor > 0 &&
                        filter_data_pipeline[0].height / (size_t)config.scale_factor > 0;
  if (do_scale) scaled = scaling(ctx_, filter_data_pipeline[0], config);
  const ScannedImageFilterData& s0 = do_scale ? scaled : filter_data_pipeline[0];
  const size_t n = s0.n(), P = s0.pixels();
  if (!gpu_size_ok(n)) { last_error = "unsupported trace length"; return THZ_EINVAL; }
  auto active = [&](const std::string& name) {
    const std::string u = uuid_of(name);
    return !u.empty() && filters_active[u];
  };
  auto* tilt = dynamic_cast<TiltCompensation*>(filter_by_name("Tilt Compensation"));
  if (active("Tilt Compensation") && tilt && (tilt->tilt_x != 0.0 || tilt->tilt_y != 0.0) && s0.dx && s0.dy) {
    last_error = "non-zero tilt is not a multiplier stage";
    return THZ_EINVAL;
  }
  // sequential f32 products in chain order, like the reference's successive in-place multiplications
  std::vector<float> m_pre(n, 1.0f), tmp(n);
  if (active("Tilt Compensation") && s0.dx && s0.dy) {
    thz_adapted_blackman(s0.time.data(), (int)n, 0.0f, 7.0f, tmp.data());
    for (size_t i = 0; i < n; ++i) m_pre[i] *= tmp[i];
  }
  if (active("Time Domain Band Pass (before FFT)")) {
    auto* g = dynamic_cast<TimeGateBase*>(filter_by_name("Time Domain Band Pass (before FFT)"));
    tmp = g->multiplier(s0.time);
    for (size_t i = 0; i < n; ++i) m_pre[i] *= tmp[i];
  }
  thz_window_multiplier((int)config.fft_window_type, s0.time.data(), (int)n, config.fft_window[0],
                        config.fft_window[1], tmp.data());
  for (size_t i = 0; i < n; ++i) m_pre[i] *= tmp[i];
  std::vector<float> band, m_post;
  if (active("Frequency Domain Band Pass"))
    band = dynamic_cast<FrequencyDomainBandPass*>(filter_by_name("Frequency Domain Band Pass"))->multiplier(s0.frequency);
  if (active("Time Domain Band Pass (after FFT)"))
    m_post = dynamic_cast<TimeGateBase*>(filter_by_name("Time Domain Band Pass (after FFT)"))->multiplier(s0.time);
  int rc = thz_plan_trace(ctx_, (int)n, m_pre.data(), band.empty() ? nullptr : band.data(),
                          m_post.empty() ? nullptr : m_post.data());
  if (rc != THZ_OK) { last_error = thz_last_error(ctx_); return rc; }
  fused_out.assign(P * n, 0.f);
  fused_img.assign(P, 0.f);
  std::vector<thz_band_plan> bands;
  int n_bands = 0;
  if (run_deconvolution && active("Deconvolution")) {
    auto* d = dynamic_cast<Deconvolution*>(filter_by_name("Deconvolution"));
    const thz_psf psf = gui_settings.psf.view();
    thz_deconv_params prm{d->n_iterations, d->n_filters, d->start_freq, d->end_freq, d->win_width};
    bands.resize((size_t)std::max(d->n_filters, 1));
    const bool has = s0.dx && s0.dy;
    const int prc = thz_deconv_plan_bands(gui_settings.psf.loaded ? &psf : nullptr, &prm, s0.time.data(), (int)n,
                                          (int)s0.width, (int)s0.height, has ? 1 : 0, has ? *s0.dx : 0.f,
                                          has ? *s0.dy : 0.f, bands.data());
    if (prc == THZ_OK) n_bands = d->n_filters;   // otherwise the reference returns the filter's input
    else if (prc == THZ_SKIP_PSF_UNSUPPORTED || prc < 0) {
      last_error = "Deconvolution: PSF extent above THZ_MAX_PSF pixels or bad parameters (library limit)";
      return prc < 0 ? prc : THZ_EINVAL;
    }
  }
  struct Cb {
    ChainDriver* self;
    static void fn(float f, void* u) { static_cast<Cb*>(u)->self->last_progress = f; }
  } cb{this};
  rc = thz_chain_host(ctx_, s0.data.data(), (int)s0.width, (int)s0.height, (int)n, n_bands ? bands.data() : nullptr,
                      n_bands, fused_out.data(), fused_img.data(),
                      reinterpret_cast<const volatile uint8_t*>(&abort_flag), &Cb::fn, &cb);
  last_progress = std::nullopt;
  if (rc != THZ_OK) { last_error = thz_last_error(ctx_); return rc; }
  return THZ_OK;
}

}  // namespace thzhost

// ---------------------------------------------------------------------------------------
// C handle over ChainDriver (used by the Python tests; a Rust host links the classes' Rust twins)
// ---------------------------------------------------------------------------------------
using namespace thzhost;

struct thz_chain {
  ChainDriver driver;
  explicit thz_chain(thz_ctx* c) : driver(c) {}
};

extern "C" {

int thz_chain_create(thz_ctx* ctx, thz_chain** out) {
  if (!ctx || !out) return THZ_EINVAL;
  *out = new thz_chain(ctx);
  return THZ_OK;
}
void thz_chain_destroy(thz_chain* ch) { delete ch; }
int thz_chain_length(const thz_chain* ch) { return ch ? (int)ch->driver.filter_chain.size() : 0; }
/* stage name at chain position i: "scaling" / "fft" / "ifft" or the filter's config().name */
const char* thz_chain_stage_name(thz_chain* ch, int i) {
  static thread_local std::string s;
  if (!ch || i < 0 || i >= (int)ch->driver.filter_chain.size()) return "";
  const std::string& id = ch->driver.filter_chain[i];
  if (id == "scaling" || id == "fft" || id == "ifft") return id.c_str();
  for (auto& kv : FilterRegistry::global().filters)
    if (kv.first == id) {
      s = kv.second->config().name;
      return s.c_str();
    }
  return "";
}
int thz_chain_set_config(thz_chain* ch, float window_lo, float window_hi, int window_type, int scale_factor) {
  if (!ch) return THZ_EINVAL;
  ch->driver.config.fft_window[0] = window_lo;
  ch->driver.config.fft_window[1] = window_hi;
  ch->driver.config.fft_window_type = (FftWindowType)window_type;
  ch->driver.config.scale_factor = scale_factor;
  return THZ_OK;
}
int thz_chain_set_psf(thz_chain* ch, const thz_psf* psf) {
  if (!ch || !psf) return THZ_EINVAL;
  ch->driver.gui_settings.psf.assign(*psf);
  return THZ_OK;
}
int thz_chain_set_param(thz_chain* ch, const char* filter_name, const char* param, double value) {
  if (!ch || !filter_name || !param) return THZ_EINVAL;
  Filter* f = ch->driver.filter_by_name(filter_name);
  return (f && f->set_param(param, value)) ? THZ_OK : THZ_EINVAL;
}
int thz_chain_get_param(thz_chain* ch, const char* filter_name, const char* param, double* value) {
  if (!ch || !filter_name || !param || !value) return THZ_EINVAL;
  Filter* f = ch->driver.filter_by_name(filter_name);
  return (f && f->get_param(param, value)) ? THZ_OK : THZ_EINVAL;
}
int thz_chain_set_active(thz_chain* ch, const char* filter_name, int active) {
  if (!ch || !filter_name) return THZ_EINVAL;
  const std::string u = ch->driver.uuid_of(filter_name);
  if (u.empty()) return THZ_EINVAL;
  ch->driver.filters_active[u] = active != 0;
  return THZ_OK;
}
int thz_chain_open(thz_chain* ch, const float* time, int n, const float* data, int width, int height, int has_dxdy,
                   float dx, float dy) {
  if (!ch || !time || !data || n < 2 || width < 0 || height < 0) return THZ_EINVAL;
  std::vector<float> t(time, time + n);
  ch->driver.open(t, data, (size_t)width, (size_t)height, has_dxdy ? std::optional<float>(dx) : std::nullopt,
                  has_dxdy ? std::optional<float>(dy) : std::nullopt);
  return THZ_OK;
}
int thz_chain_run(thz_chain* ch, int start_idx, int run_deconvolution) {
  if (!ch) return THZ_EINVAL;
  return ch->driver.run((size_t)std::max(start_idx, 1), run_deconvolution != 0);
}
int thz_chain_run_fused(thz_chain* ch, int run_deconvolution) {
  if (!ch) return THZ_EINVAL;
  return ch->driver.run_fused(run_deconvolution != 0);
}
void thz_chain_abort(thz_chain* ch, int value) {
  if (ch) ch->driver.abort_flag.store(value != 0);
}
/* borrow the arrays of pipeline slot `slot` (0 = loaded scan; i + 1 = output of chain stage i) */
int thz_chain_slot(thz_chain* ch, int slot, const float** data, const float** fft, const float** amp,
                   const float** phase, const float** img, const float** avg_fft, const float** avg_amp,
                   const float** avg_phase, int* n, int* f) {
  if (!ch || slot < 0 || slot >= (int)ch->driver.filter_data_pipeline.size()) return THZ_EINVAL;
  const ScannedImageFilterData& s = ch->driver.filter_data_pipeline[slot];
  if (data) *data = s.data.data();
  if (fft) *fft = reinterpret_cast<const float*>(s.fft.data());
  if (amp) *amp = s.amplitudes.data();
  if (phase) *phase = s.phases.data();
  if (img) *img = s.img.data();
  if (avg_fft) *avg_fft = reinterpret_cast<const float*>(s.avg_fft.data());
  if (avg_amp) *avg_amp = s.avg_signal_fft.data();
  if (avg_phase) *avg_phase = s.avg_phase_fft.data();
  if (n) *n = (int)s.n();
  if (f) *f = (int)s.f();
  return THZ_OK;
}
int thz_chain_fused_result(thz_chain* ch, const float** data, const float** img) {
  if (!ch) return THZ_EINVAL;
  if (data) *data = ch->driver.fused_out.data();
  if (img) *img = ch->driver.fused_img.data();
  return THZ_OK;
}
double thz_chain_filter_ms(thz_chain* ch, const char* filter_name) {
  if (!ch || !filter_name) return -1.0;
  const std::string u = ch->driver.uuid_of(filter_name);
  if (u.empty()) return -1.0;
  return ch->driver.filter_computation_time[u].count() * 1e3;
}

}  // extern "C"
