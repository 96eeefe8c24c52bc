// thz_deconv_plan.cpp -- host part of the PSF deconvolution filter: Kaiser FIR bank design,
// PSF model evaluation and per-band planning.  Tiny, runs once per filter application, and
// is restated in the reference's own precision (FIR design in f64, PSF model in f32).
//   kaiser_atten/beta, i0, sinc, firwin_kaiser_*, bandpass_kaiser, create_filter_bank
//                                              ... src/filters/deconvolution.rs:30-211
//   CubicSplineCoeffs / HybridFit evaluators ... src/filters/psf.rs:26-179
//   gaussian, create_psf_2d .................... src/filters/psf.rs:228-332
//   band loop of Deconvolution::filter ......... src/filters/deconvolution.rs:781-971
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/thzgpu.h"

namespace {

// ---- FIR design (f64) ---------------------------------------------------------------
double kaiser_atten(int ntaps, double width_ratio) {
  const double a = 2.285 * ((double)ntaps - 1.0) * M_PI * width_ratio + 7.95;
  return std::max(a, 0.0);
}
double kaiser_beta(double atten) {
  if (atten > 50.0) return 0.1102 * (atten - 8.7);
  if (atten >= 21.0) return 0.5842 * pow(atten - 21.0, 0.4) + 0.07886 * (atten - 21.0);
  return 0.0;
}
double bessel_i0(double x) {
  double sum = 1.0, term = 1.0;
  const double xh = (x / 2.0) * (x / 2.0);
  for (int k = 1; k < 50; ++k) {
    term *= xh / (double)(k * k);
    sum += term;
    if (term < 1e-12 * sum) break;
  }
  return sum;
}
double sinc(double x) { return fabs(x) < 1e-10 ? 1.0 : sin(x) / x; }
double kaiser_coeff(int n, int n_taps, double beta) {
  if (n == 0 || n == n_taps - 1) return 0.0;
  const double arg = 2.0 * (double)n / ((double)n_taps - 1.0) - 1.0;
  return bessel_i0(beta * sqrt(1.0 - arg * arg)) / bessel_i0(beta);
}
std::vector<double> lowpass(int n_taps, double cutoff_hz, double beta, double fs) {
  const int adj = (n_taps % 2 == 0) ? n_taps - 1 : n_taps;
  const double mid = (double)(adj - 1) / 2.0;
  const double cutoff = cutoff_hz / fs;
  std::vector<double> f(adj);
  for (int n = 0; n < adj; ++n) f[n] = sinc(2.0 * M_PI * cutoff * ((double)n - mid)) * kaiser_coeff(n, adj, beta);
  double sum = 0.0;
  for (double v : f) sum += v;
  if (fabs(sum) > 1e-10)
    for (double& v : f) v /= sum;
  if (n_taps % 2 == 0) f.push_back(0.0);
  return f;
}
std::vector<double> highpass(int n_taps, double cutoff_hz, double beta, double fs) {
  const int adj = (n_taps % 2 == 0) ? n_taps - 1 : n_taps;
  const double mid = (double)(adj - 1) / 2.0;
  std::vector<double> f = lowpass(adj, cutoff_hz, beta, fs);
  for (int i = 0; i < (int)f.size(); ++i) f[i] = (i == (int)mid) ? 1.0 - f[i] : -f[i];
  if (n_taps % 2 == 0) f.push_back(0.0);
  return f;
}
std::vector<double> bandpass_kaiser(int ntaps, double lowcut, double highcut, double fs, double width) {
  const double beta = kaiser_beta(kaiser_atten(ntaps, width / (0.5 * fs)));
  if (lowcut <= 0.0) return lowpass(ntaps, highcut, beta, fs);
  if (highcut >= 0.5 * fs) return highpass(ntaps, lowcut, beta, fs);
  std::vector<double> lo = highpass(ntaps, lowcut, beta, fs), hi = highpass(ntaps, highcut, beta, fs);
  for (size_t i = 0; i < lo.size(); ++i) lo[i] -= hi[i];
  return lo;
}

// ---- PSF model (f32) ------------------------------------------------------------------
float poly(const thz_spline& s, int i, float dx) {
  return s.coeff_a[i] + s.coeff_b[i] * dx + s.coeff_c[i] * dx * dx + s.coeff_d[i] * dx * dx * dx;
}
int segment(const thz_spline& s, float x) {
  int left = 0, right = s.n - 1;
  while (right - left > 1) {
    const int mid = (left + right) / 2;
    if (s.knots[mid] > x) right = mid;
    else left = mid;
  }
  return left;
}
// src/filters/psf.rs:26-80
float spline_eval(const thz_spline& s, float x) {
  const int n = s.n;
  if (n == 0) return 0.0f;
  if (x < s.knots[0]) {
    const float dx = x - s.knots[0];
    return std::max(s.coeff_a[0] + s.coeff_b[0] * dx, 1e-6f);
  }
  if (x > s.knots[n - 1]) {
    const int i = n - 2;
    const float dxe = s.knots[n - 1] - s.knots[i];
    const float y_end = poly(s, i, dxe);
    const float slope = s.coeff_b[i] + 2.0f * s.coeff_c[i] * dxe + 3.0f * s.coeff_d[i] * dxe * dxe;
    return std::max(y_end + slope * (x - s.knots[n - 1]), 1e-6f);
  }
  const int l = segment(s, x);
  return poly(s, l, x - s.knots[l]);
}
// src/filters/psf.rs:83-117
float spline_eval_const(const thz_spline& s, float x) {
  const int n = s.n;
  if (n == 0) return 0.0f;
  if (x < s.knots[0]) return s.values[0];
  if (x > s.knots[n - 1]) return s.values[n - 1];
  const int l = segment(s, x);
  return poly(s, l, x - s.knots[l]);
}
// src/filters/psf.rs:134-179
float hybrid_correction(const thz_hybrid_fit& h, float f) {
  const thz_spline& c = h.correction;
  const int n = c.n;
  if (n == 0) return 0.0f;
  const float f_min = c.knots[0], f_max = c.knots[n - 1];
  if (f >= f_min && f <= f_max) return spline_eval(c, f);
  if (f < f_min) {
    const float dx = f - f_min;
    const float max_slope = h.base_a / (f * f);
    return c.coeff_a[0] + std::min(c.coeff_b[0], max_slope) * dx;
  }
  const int i = n - 2;
  const float dxe = c.knots[n - 1] - c.knots[i];
  const float y_end = poly(c, i, dxe);
  const float slope_end = c.coeff_b[i] + 2.0f * c.coeff_c[i] * dxe + 3.0f * c.coeff_d[i] * dxe * dxe;
  const float max_slope = h.base_a / (f * f);
  return y_end + std::min(slope_end, max_slope) * (f - c.knots[n - 1]);
}
// src/filters/psf.rs:122-131
float hybrid_eval(const thz_hybrid_fit& h, float f) {
  const float base = h.base_a / f + h.base_b;
  return std::max(base + hybrid_correction(h, f), 1e-6f);
}
// src/filters/psf.rs:326-332
float gaussian(float xi, float x0, float w) {
  const float kPi = 3.14159265358979323846f;
  const float d = xi - x0;
  return sqrtf(2.0f / kPi) * expf(-2.0f * (d * d) / (w * w)) / w;
}

}  // namespace

extern "C" {

int thz_fir_bank(int n_filters, double start_freq, double end_freq, double win_width, float t0, float t1,
                 float* filters, float* center_freqs) {
  if (n_filters < 1 || n_filters > THZ_MAX_BANDS || !filters || !center_freqs) return THZ_EINVAL;
  const int ntaps = THZ_FIR_TAPS;
  const double dt = (double)(t1 - t0);   // f32 subtraction, then widened (deconvolution.rs:170)
  const double fs = 1.0 / dt;
  const double log_start = log(start_freq), log_end = log(end_freq);
  const double log_step = (log_end - log_start) / (double)(n_filters - 1);
  for (int i = 0; i < n_filters; ++i) center_freqs[i] = (float)exp(log_start + (double)i * log_step);
  for (int i = 0; i < n_filters; ++i) {
    const double cf = (double)center_freqs[i];
    const double lowcut = (i == 0) ? 0.0 : sqrt((double)center_freqs[i - 1] * cf);
    const double highcut = (i == n_filters - 1) ? 0.5 * fs : sqrt(cf * (double)center_freqs[i + 1]);
    std::vector<double> h = bandpass_kaiser(ntaps, lowcut, highcut, fs, win_width);
    for (int j = 0; j < ntaps; ++j) filters[(size_t)i * ntaps + j] = (j < (int)h.size()) ? (float)h[j] : 0.0f;
  }
  return THZ_OK;
}

float thz_hybrid_eval(const thz_hybrid_fit* fit, float f) { return fit ? hybrid_eval(*fit, f) : 0.0f; }
float thz_spline_eval_const_extrap(const thz_spline* s, float f) { return s ? spline_eval_const(*s, f) : 0.0f; }

int thz_deconv_plan_bands(const thz_psf* psf, const thz_deconv_params* prm, const float* time, int n,
                          int img_rows, int img_cols, int has_dxdy, float dx, float dy, thz_band_plan* bands) {
  if (!prm || !time || n < 2 || !bands) return THZ_EINVAL;
  if (!has_dxdy) return THZ_SKIP_NO_DXDY;
  if (!psf || psf->wx_fit.correction.n == 0) return THZ_SKIP_NO_PSF;
  const int kMinImage = 16;
  if (img_rows < kMinImage || img_cols < kMinImage) return THZ_SKIP_TOO_SMALL;
  const int B = prm->n_filters;
  if (B < 1 || B > THZ_MAX_BANDS) return THZ_EINVAL;
  std::vector<float> filters((size_t)B * THZ_FIR_TAPS), centers(B);
  int rc = thz_fir_bank(B, (double)prm->start_freq, (double)prm->end_freq, (double)prm->win_width, time[0], time[1],
                        filters.data(), centers.data());
  if (rc != THZ_OK) return rc;
  float wx_min = INFINITY, wx_max = -INFINITY, wy_min = INFINITY, wy_max = -INFINITY;
  for (int i = 0; i < B; ++i) {
    const float wx = hybrid_eval(psf->wx_fit, centers[i]), wy = hybrid_eval(psf->wy_fit, centers[i]);
    wx_min = std::min(wx_min, wx); wx_max = std::max(wx_max, wx);
    wy_min = std::min(wy_min, wy); wy_max = std::max(wy_max, wy);
  }
  const float w_min = std::min(wx_min, wy_min), w_max = std::max(wx_max, wy_max);
  const long max_w_x = std::max((long)ceilf(wx_max / dx) * 2 + 1, 3L);
  const long max_w_y = std::max((long)ceilf(wy_max / dy) * 2 + 1, 3L);
  if (max_w_x >= img_cols || max_w_y >= img_rows) return THZ_SKIP_PSF_TOO_LARGE;

  for (int i = 0; i < B; ++i) {
    thz_band_plan& b = bands[i];
    memset(&b, 0, sizeof b);
    const float cf = centers[i];
    b.center_freq = cf;
    b.wx = hybrid_eval(psf->wx_fit, cf);
    b.wy = hybrid_eval(psf->wy_fit, cf);
    b.x0 = spline_eval_const(psf->x0_spline, cf);
    b.y0 = spline_eval_const(psf->y0_spline, cf);
    // support: max(3 (w + |c|), 2.5 mm), snapped to the pixel grid, clamped to the image (:920-951)
    float rx = std::max((b.wx + fabsf(b.x0)) * 3.0f, 2.5f);
    float ry = std::max((b.wy + fabsf(b.y0)) * 3.0f, 2.5f);
    rx = floorf(rx / dx) * dx + dx;
    ry = floorf(ry / dy) * dy + dy;
    const float max_allowed_x = ((float)img_cols - 2.0f) * dx / 2.0f;
    const float max_allowed_y = ((float)img_rows - 2.0f) * dy / 2.0f;
    const float crx = std::min(rx, max_allowed_x), cry = std::min(ry, max_allowed_y);
    const int hx = (int)floorf(crx / dx), hy = (int)floorf(cry / dy);
    std::vector<float> x(2 * hx + 1), y(2 * hy + 1), gx(2 * hx + 1), gy(2 * hy + 1);
    for (int v = -hx; v <= hx; ++v) x[v + hx] = (float)v * dx;
    for (int v = -hy; v <= hy; ++v) y[v + hy] = (float)v * dy;
    float gx_max = -INFINITY, gy_max = -INFINITY, x_maxv = -INFINITY, y_maxv = -INFINITY;
    for (size_t k = 0; k < x.size(); ++k) {
      gx[k] = gaussian(x[k], b.x0, b.wx);
      gx_max = std::max(gx_max, gx[k]);
      x_maxv = std::max(x_maxv, x[k]);
    }
    for (size_t k = 0; k < y.size(); ++k) {
      gy[k] = gaussian(y[k], b.y0, b.wy);
      gy_max = std::max(gy_max, gy[k]);
      y_maxv = std::max(y_maxv, y[k]);
    }
    // create_psf_2d (psf.rs:228-313): factors normalised to max 1; the grid is v*dx for
    // |v| <= floor(max x in mm) -- the extent in PIXELS is 2 floor(range_mm) + 1 -- and the
    // linear interpolator is only queried at its own knots (a lookup; zero outside the support)
    const int vx = (int)floorf(x_maxv), vy = (int)floorf(y_maxv);
    b.kx = 2 * vx + 1;
    b.ky = 2 * vy + 1;
    if (b.kx > THZ_MAX_PSF || b.ky > THZ_MAX_PSF) return THZ_SKIP_PSF_UNSUPPORTED;   // library limit, not a reference skip
    for (int v = -vx; v <= vx; ++v) b.psf_x[v + vx] = (v >= -hx && v <= hx) ? gx[v + hx] / gx_max : 0.0f;
    for (int v = -vy; v <= vy; ++v) b.psf_y[v + vy] = (v >= -hy && v <= hy) ? gy[v + hy] / gy_max : 0.0f;
    b.n_iter = (int)floorf((b.wx - w_min) / (w_max - w_min) * ((float)prm->n_iterations - 1.0f) + 1.0f);
    b.direct = (b.kx * b.ky <= 256) ? 1 : 0;   // convolve2d's branch (deconvolution.rs:484)
    memcpy(b.fir, &filters[(size_t)i * THZ_FIR_TAPS], sizeof(float) * THZ_FIR_TAPS);
  }
  return THZ_OK;
}

}  // extern "C"
