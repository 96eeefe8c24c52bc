// thz_windows.cpp -- host-side, pixel-independent multiplier vectors of the filter chain.
//
// Every time-domain filter on the default path and the frequency band-pass reduce to a
// vector that depends only on the axis and the filter parameters.  They are
// computed here once per run, in f32 exactly as the reference computes them per trace, and
// handed to the kernels through thz_plan_trace().
//   blackman_window / apply_adapted_blackman_window ... src/math_tools.rs:81-122
//   normalize_time, hamming/hanning/blackman/flat-top . src/math_tools.rs:131-198
//   frequency axis ................................... src/io.rs:614-620
//   time gate indices + taper ........................ src/filters/band_pass_td_before_fft.rs:134-174
//   band-pass indices + taper ........................ src/filters/band_pass_fd.rs:135-212
#include <math.h>

#include <algorithm>
#include <vector>

#include "../../include/thzgpu.h"

namespace {

const float kPi = 3.14159265358979323846f;  // std::f32::consts::PI

// src/math_tools.rs:81-90
float blackman_window(float n, float m) {
  const float res = 0.42f - 0.5f * cosf(2.0f * kPi * n / m) + 0.08f * cosf(4.0f * kPi * n / m);
  if (isnan(res)) return 1.0f;
  return std::min(std::max(res, 0.0f), 1.0f);
}

// src/math_tools.rs:102-122 applied to a vector of ones
void adapted_blackman(const float* axis, int n, float lo, float hi, float* mult) {
  for (int i = 0; i < n; ++i) mult[i] = 1.0f;
  if (n == 0) return;
  const float a0 = axis[0], al = axis[n - 1];
  for (int i = 0; i < n; ++i) {
    const float t = axis[i];
    if (t <= lo + a0) {
      mult[i] = blackman_window(t - a0, 2.0f * lo);
    } else if (t >= al - hi) {
      mult[i] = blackman_window(t - (al - hi * 2.0f), 2.0f * hi);
    }
  }
}

}  // namespace

extern "C" {

int thz_frequency_axis(const float* time, int n, float* freq) {
  if (!time || !freq || n < 2) return THZ_EINVAL;
  const float rng = time[n - 1] - time[0];
  for (int i = 0; i < n / 2 + 1; ++i) freq[i] = (float)i / rng;
  return THZ_OK;
}

int thz_adapted_blackman(const float* axis, int n, float lo, float hi, float* mult) {
  if ((!axis || !mult) && n > 0) return THZ_EINVAL;
  if (n < 0) return THZ_EINVAL;
  adapted_blackman(axis, n, lo, hi, mult);
  return THZ_OK;
}

int thz_window_multiplier(int window_type, const float* time, int n, float lo, float hi, float* mult) {
  if (!time || !mult || n <= 0) return THZ_EINVAL;
  if (window_type == THZ_WINDOW_ADAPTED_BLACKMAN) {
    adapted_blackman(time, n, lo, hi, mult);
    return THZ_OK;
  }
  // normalize_time (src/math_tools.rs:131-135)
  float mn = INFINITY, mx = -INFINITY;
  for (int i = 0; i < n; ++i) {
    mn = std::min(mn, time[i]);
    mx = std::max(mx, time[i]);
  }
  for (int i = 0; i < n; ++i) {
    const float t = (time[i] - mn) / (mx - mn);
    switch (window_type) {
      case THZ_WINDOW_BLACKMAN:
        mult[i] = 0.42f - 0.5f * cosf(2.0f * kPi * t) + 0.08f * cosf(4.0f * kPi * t);
        break;
      case THZ_WINDOW_HANNING:
        mult[i] = 0.5f * (1.0f - cosf(2.0f * kPi * t));
        break;
      case THZ_WINDOW_HAMMING:
        mult[i] = 0.54f - 0.46f * cosf(2.0f * kPi * t);
        break;
      case THZ_WINDOW_FLAT_TOP:
        mult[i] = 1.0f - 1.93f * cosf(2.0f * kPi * t) + 1.29f * cosf(4.0f * kPi * t) -
                  0.388f * cosf(6.0f * kPi * t) + 0.028f * cosf(8.0f * kPi * t);
        break;
      default:
        return THZ_EINVAL;
    }
  }
  return THZ_OK;
}

int thz_time_gate_multiplier(const float* time, int n, double* low, double* high, double window_width,
                             float* mult, int* lower_out, int* upper_out) {
  if (!time || !mult || !low || !high || n <= 0) return THZ_EINVAL;
  const float min_time = time[0], max_time = time[n - 1];
  *low = std::max(*low, (double)min_time);     // the filter clamps its own fields (:134-138)
  *high = std::min(*high, (double)max_time);
  int lower = 0;
  {
    const float lo32 = (float)*low;
    int i = 0;
    while (i < n && !(time[i] >= lo32)) ++i;
    lower = (i < n) ? i : 0;
  }
  int upper;
  {
    const float hi32 = (float)*high;
    int i = 0;
    while (i < n && !(time[i] >= hi32)) ++i;
    upper = (i < n) ? i : std::max(n - 1, 0);
  }
  upper = std::min(std::max(upper, lower + 1), n);
  for (int i = 0; i < n; ++i) mult[i] = 0.0f;
  adapted_blackman(time + lower, upper - lower, (float)window_width, (float)window_width, mult + lower);
  if (lower_out) *lower_out = lower;
  if (upper_out) *upper_out = upper;
  return THZ_OK;
}

int thz_band_pass_multiplier(const float* freq, int f, double low, double high, double window_width, float* mult,
                             int* lower_out, int* upper_out) {
  if (!freq || !mult || f <= 0) return THZ_EINVAL;
  const float safe_low = (float)std::max(low, 0.0);
  const float safe_high = (float)std::min(high, (double)freq[f - 1]);
  int lower = 0;
  {
    int i = 0;
    while (i < f && !(freq[i] >= safe_low)) ++i;
    lower = (i < f) ? i : 0;
  }
  int upper = f;
  {
    int i = f - 1;
    while (i >= 0 && !(freq[i] <= safe_high)) --i;
    upper = (i >= 0) ? i + 1 : f;
  }
  for (int i = 0; i < f; ++i) mult[i] = 0.0f;
  if (upper > lower)
    adapted_blackman(freq + lower, upper - lower, (float)window_width, (float)window_width, mult + lower);
  if (lower_out) *lower_out = lower;
  if (upper_out) *upper_out = upper;
  return THZ_OK;
}

/* `calculate_optical_properties` (src/math_tools.rs:663-701): refractive index, absorption and extinction
 * coefficient per bin from a sample and a reference spectrum (selected pixel or ROI mean; F values). */
int thz_optical_properties(const float* sample_amp, const float* sample_phase, const float* ref_amp,
                           const float* ref_phase, const float* freqs, int f, float thickness, float* n_out,
                           float* alpha_out, float* kappa_out) {
  if (!sample_amp || !sample_phase || !ref_amp || !ref_phase || !freqs || f < 0) return THZ_EINVAL;
  const float kC = 2.99792458e8f;
  for (int i = 0; i < f; ++i) {
    const float frequency_hz = freqs[i] * 1.0e12f;
    const float delta_phi = sample_phase[i] - ref_phase[i];
    const float omega = 2.0f * kPi * frequency_hz;
    const float n = 1.0f + kC * delta_phi / (omega * thickness);
    const float amp = std::max(sample_amp[i], 1e-12f), amp_ref = std::max(ref_amp[i], 1e-12f);
    const float n_safe = std::max(n, 1e-6f);
    const float np1 = n_safe + 1.0f;
    const float alpha = -2.0f / thickness * logf((np1 * np1) / (4.0f * n_safe) * amp / amp_ref);
    const float kappa = alpha * kC / (4.0f * kPi * frequency_hz);
    if (n_out) n_out[i] = n;
    if (alpha_out) alpha_out[i] = alpha;
    if (kappa_out) kappa_out[i] = kappa;
  }
  return THZ_OK;
}

/* Host part of `TiltCompensation::filter` (src/filters/tilt_compensation.rs:104-168): number of extension
 * steps, extended time axis (front linspace | original | back linspace) and the per-pixel insert index
 * max(num_steps + floor(delta / dt), 0), in the reference's mixed f32 / f64 arithmetic.  time_ext must hold
 * n + 2 * (*num_steps) floats -- call once with time_ext = insert = NULL to get num_steps. */
int thz_tilt_plan(const float* time, int n, int width, int height, float dx, float dy, double tilt_x, double tilt_y,
                  int* num_steps_out, float* time_ext, int* insert) {
  if (!time || n < 1 || width < 1 || height < 1 || !num_steps_out) return THZ_EINVAL;
  const float time_shift_x = (float)tilt_x / 180.0f * kPi;
  const float time_shift_y = (float)tilt_y / 180.0f * kPi;
  const float center_x = (float)width / 2.0f * dx, center_y = (float)height / 2.0f * dy;
  const double c = 0.299792458;
  const float dt = 0.05f;
  const float max_offset_x = (float)((double)center_x * (double)fabsf(time_shift_x) / c);
  const float max_offset_y = (float)((double)center_y * (double)fabsf(time_shift_y) / c);
  float extension = (max_offset_x + max_offset_y) / dt;
  extension = floorf(extension) * dt;
  const int num_steps = (int)roundf(extension / dt);
  *num_steps_out = num_steps;
  if (time_ext) {
    const float first = time[0], last = time[n - 1];
    // ndarray `Array1::linspace(a, b, k)`: a + i * (b - a) / (k - 1)
    auto linspace = [](float a, float b, int k, float* out) {
      if (k == 1) out[0] = a;
      const float step = (k > 1) ? (b - a) / (float)(k - 1) : 0.f;
      for (int i = 0; i < k; ++i) out[i] = a + step * (float)i;
    };
    linspace(first - extension, first - dt, num_steps, time_ext);
    for (int i = 0; i < n; ++i) time_ext[num_steps + i] = time[i];
    linspace(last + dt, last + extension, num_steps, time_ext + num_steps + n);
  }
  if (insert) {
    for (int i = 0; i < width; ++i) {
      const float x_off = (float)((double)(((float)i - (float)width / 2.0f) * dx) * (double)time_shift_x / c);
      for (int j = 0; j < height; ++j) {
        const float y_off = (float)((double)(((float)j - (float)height / 2.0f) * dy) * (double)time_shift_y / c);
        const float delta = x_off + y_off;
        const long steps = (long)floorf(delta / dt);
        insert[(size_t)i * height + j] = (int)std::max((long)num_steps + steps, 0L);
      }
    }
  }
  return THZ_OK;
}

}  // extern "C"
