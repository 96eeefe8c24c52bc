// thz_host.hpp -- C++ mirror of the reference's plugin API for the filter-chain hot path.
//
// The reference is Rust; no Rust toolchain exists in this build environment, so the host side
// above the C ABI is written in C++ with the same names, argument meaning and error behaviour
// (INTEGRATION.md shows the equivalent Rust shim a maintainer would add):
//   ScannedImageFilterData ....... src/data_container.rs:109-162
//   ConfigContainer, FftWindowType  src/config.rs:171-212
//   Filter, FilterConfig, FilterDomain, FilterRegistry, FILTER_REGISTRY
//                                   src/filters/filter.rs:96-221, 232-262, 319-338, 448-452
//   #[register_filter] ............ filter_macros/src/lib.rs:45-69  ->  THZ_REGISTER_FILTER
//   scaling / fft / ifft .......... src/math_tools.rs:242, 330, 418
//   chain assembly ................ src/main.rs:194-268
//   chain driver loop ............. src/data_thread.rs:1090-1228, 1288-1307
// All arithmetic is done by libthzgpu (include/thzgpu.h); nothing here computes on the CPU
// except the pixel-independent multiplier vectors (thz_windows.cpp) and the band planner.
#pragma once
#include <atomic>
#include <chrono>
#include <complex>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <vector>

#include "../../../include/thzgpu.h"

namespace thzhost {

enum class FftWindowType { AdaptedBlackman = 0, Blackman = 1, Hanning = 2, Hamming = 3, FlatTop = 4 };

// src/config.rs:171-212
struct ConfigContainer {
  float fft_window[2] = {1.0f, 7.0f};
  FftWindowType fft_window_type = FftWindowType::AdaptedBlackman;
  int scale_factor = 1;
  bool avg_in_fourier_space = false;
};

// the part of `GuiSettingsContainer` the filters read (the loaded PSF, src/filters/psf.rs:202-207)
struct PsfStorage {
  std::vector<float> arrays[4][6];   // wx corr, wy corr, x0, y0 x {knots, values, a, b, c, d}
  float base_a[2] = {0, 0}, base_b[2] = {0, 0};
  bool loaded = false;
  thz_psf view() const;
  void assign(const thz_psf& p);
};
struct GuiSettingsContainer {
  PsfStorage psf;
};

// src/data_container.rs:109-162 (fields on the hot path)
struct ScannedImageFilterData {
  std::vector<float> time, frequency;
  std::vector<float> data;                        // [width][height][N]
  std::vector<std::complex<float>> fft;           // [width][height][F]
  std::vector<float> amplitudes, phases;          // [width][height][F]
  std::vector<float> img;                         // [width][height]
  std::vector<std::complex<float>> avg_fft;
  std::vector<float> avg_signal_fft, avg_phase_fft;
  std::optional<float> dx, dy;
  size_t width = 0, height = 0;
  bool has_plan = false;                          // r2c / c2r present
  size_t n() const { return time.size(); }
  size_t f() const { return frequency.size(); }
  size_t pixels() const { return width * height; }
};

enum class FilterDomain { TimeBeforeFFTPrioFirst, TimeBeforeFFT, Frequency, TimeAfterFFT, TimeAfterFFTPrioLast };

struct FilterConfig {
  std::string name, description;
  FilterDomain domain;
};

using ProgressLock = std::function<void(std::optional<float>)>;   // Arc<RwLock<Option<f32>>>

struct FilterEnv {
  thz_ctx* ctx = nullptr;   // process-global in the Rust shim: filters are cloned on every update
  std::string last_error;   // what the Rust shim would `log::error!` (set when a filter passes its input through
                            // for a reason the reference does not have, e.g. a PSF above THZ_MAX_PSF)
};
// std::atomic<bool> has the one-byte layout of Rust's AtomicBool; the C ABI polls it as a byte
static_assert(sizeof(std::atomic<bool>) == 1, "abort flag must be one byte");

// src/filters/filter.rs:96-221 (ui() is GUI code and stays on the reference side)
class Filter {
 public:
  virtual ~Filter() = default;
  virtual void reset(const std::vector<float>& time, const size_t shape[3]) = 0;
  virtual void show_data(const ScannedImageFilterData&) {}
  virtual FilterConfig config() const = 0;
  virtual ScannedImageFilterData filter(const ScannedImageFilterData& input, GuiSettingsContainer& gui_settings,
                                        ProgressLock& progress_lock, const std::atomic<bool>& abort_flag) = 0;
  virtual std::unique_ptr<Filter> clone_box() const = 0;
  // parameter access for the C test driver (the reference edits public struct fields from its ui())
  virtual bool set_param(const std::string& name, double value) = 0;
  virtual bool get_param(const std::string& name, double* value) const = 0;
  FilterEnv env;
};

// src/filters/filter.rs:319-338, 448-452
class FilterRegistry {
 public:
  static FilterRegistry& global();   // FILTER_REGISTRY
  template <class F> static void register_filter() { global().add(std::make_unique<F>()); }
  std::map<std::string, std::unique_ptr<Filter>> filters;   // uuid -> instance
  Filter* get_filter(const std::string& name);
 private:
  void add(std::unique_ptr<Filter> f);
};

// #[register_filter] (filter_macros/src/lib.rs:45-69): a static initialiser instead of #[ctor]
#define THZ_REGISTER_FILTER(T) \
  static const bool thz_registered_##T = (::thzhost::FilterRegistry::register_filter<T>(), true)

// built-in stages, same signatures as src/math_tools.rs:242, 330, 418
ScannedImageFilterData scaling(thz_ctx* ctx, const ScannedImageFilterData& input, const ConfigContainer& config);
ScannedImageFilterData fft(thz_ctx* ctx, const ScannedImageFilterData& input, const ConfigContainer& config);
ScannedImageFilterData ifft(thz_ctx* ctx, const ScannedImageFilterData& input, const ConfigContainer& config);

// the five shipped filters
class TiltCompensation;
class TimeDomainBandPassBeforeFFT;
class FrequencyDomainBandPass;
class TimeDomainBandPassAfterFFT;
class Deconvolution;

// chain assembly (src/main.rs:194-268) + driver loop (src/data_thread.rs:1023-1316)
class ChainDriver {
 public:
  explicit ChainDriver(thz_ctx* ctx);
  std::vector<std::string> filter_chain;                     // "scaling", uuids, "fft", uuids, "ifft", uuids
  std::map<std::string, size_t> filter_uuid_to_index;        // id -> slot (1-based; slot 0 = loaded scan)
  std::map<std::string, bool> filters_active;
  std::map<std::string, std::chrono::duration<double>> filter_computation_time;
  std::vector<ScannedImageFilterData> filter_data_pipeline;  // one slot per stage + slot 0
  size_t fft_index = 0, ifft_index = 0, scaling_index = 0;
  ConfigContainer config;
  GuiSettingsContainer gui_settings;
  std::atomic<bool> abort_flag{false};
  std::optional<float> last_progress;

  // ConfigCommand::OpenFile: slot 0 <- scan, reset every filter (data_thread.rs:1027-1060)
  void open(const std::vector<float>& time, const float* data, size_t width, size_t height,
            std::optional<float> dx, std::optional<float> dy);
  // UpdateType::Filter(start_idx): run filter_chain[start_idx-1 ..] (data_thread.rs:1090-1316)
  int run(size_t start_idx, bool run_deconvolution);
  // The same chain as ONE fused kernel when every active stage between "scaling" and the last time
  // gate is a multiplier (the default chain): only the last slot and the intensity image are produced.
  int run_fused(bool run_deconvolution);
  Filter* filter_by_name(const std::string& name);
  std::string uuid_of(const std::string& name) const;
  std::string last_error;
  std::vector<float> fused_out, fused_img;

 private:
  thz_ctx* ctx_;
  std::map<std::string, std::unique_ptr<Filter>> filters_;   // working clones (data_thread.rs:1064-1078)
};

}  // namespace thzhost
