/* thz_oracle_c.c -- C / pthreads twin of the oracle's default filter chain.
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/thz_oracle.py): it is the timed "port" of
 * the reference's multithreaded CPU chain for bench.py's cpu_baseline and --impl reference legs and
 * is cross-checked against the numpy oracle in tests/test_oracle_c.py.  The product never links it.
 *
 * It restates, stage by stage and with the reference's own threading, what the Rust data thread
 * does for UpdateType::Filter(1) on the default chain (src/data_thread.rs:1090-1316):
 *   every stage starts with a deep clone of ScannedImageFilterData (data, fft, amplitudes, phases)
 *       -- src/math_tools.rs:331,419; band_pass_fd.rs:129; band_pass_td_before_fft.rs:131
 *   scaling (s = 1: clone) ........................ src/math_tools.rs:242-310
 *   tilt taper at 0 deg, serial over pixels ....... src/filters/tilt_compensation.rs:170-199
 *   time gates, serial over pixels ................ src/filters/band_pass_td_before_fft.rs:155-174
 *   fft: rayon over axis 0, per trace window + r2c + |s| + unwrap(arg s) ... src/math_tools.rs:330-398
 *   FD band-pass: rows in parallel but serialised by two mutexes, then a serial zero-pad copy
 *                                                   src/filters/band_pass_fd.rs:155-212
 *   ifft: pixel means (serial), rayon over axis 0, c2r / N ................ src/math_tools.rs:418-571
 *   intensity image: rayon over rows ............... src/data_thread.rs:1288-1307
 * The FFT is a plain iterative radix-2 complex transform used through the usual N/2 real-FFT split
 * (realfft / rustfft are not available here; this is slower than their SIMD kernels, which is why the
 * number is labelled "port").  Power-of-two N only.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

/* minimal fork-join "rayon": `threads` workers pull row indices from an atomic counter */
typedef void (*row_fn)(int row, void* ctx);
typedef struct { row_fn fn; void* ctx; int rows; atomic_int next; } pool_job;
static void* pool_worker(void* arg) {
  pool_job* j = (pool_job*)arg;
  for (;;) {
    const int r = atomic_fetch_add(&j->next, 1);
    if (r >= j->rows) break;
    j->fn(r, j->ctx);
  }
  return NULL;
}
static int g_threads = 1;
static void parallel_rows(row_fn fn, void* ctx, int rows) {
  pool_job j;
  j.fn = fn; j.ctx = ctx; j.rows = rows;
  atomic_init(&j.next, 0);
  int nt = g_threads < rows ? g_threads : rows;
  if (nt <= 1) { pool_worker(&j); return; }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nt);
  for (int i = 0; i < nt; ++i) pthread_create(&th[i], NULL, pool_worker, &j);
  for (int i = 0; i < nt; ++i) pthread_join(th[i], NULL);
  free(th);
}

typedef struct { float re, im; } cf;

typedef struct {
  int n;        /* real length */
  int h;        /* n / 2 */
  cf* w;        /* twiddles exp(-2 pi i k / h), k < h/2 */
  cf* wr;       /* real-split twiddles exp(-2 pi i k / n), k <= h/2 .. h */
  int* rev;     /* bit reversal for size h */
} plan_t;

static plan_t* plan_new(int n) {
  plan_t* p = (plan_t*)malloc(sizeof(plan_t));
  p->n = n;
  p->h = n / 2;
  p->w = (cf*)malloc(sizeof(cf) * (size_t)(p->h / 2 + 1));
  p->wr = (cf*)malloc(sizeof(cf) * (size_t)(p->h + 1));
  p->rev = (int*)malloc(sizeof(int) * (size_t)p->h);
  for (int k = 0; k < p->h / 2 + 1; ++k) {
    const double a = -2.0 * M_PI * k / p->h;
    p->w[k].re = (float)cos(a);
    p->w[k].im = (float)sin(a);
  }
  for (int k = 0; k <= p->h; ++k) {
    const double a = -2.0 * M_PI * k / n;
    p->wr[k].re = (float)cos(a);
    p->wr[k].im = (float)sin(a);
  }
  int bits = 0;
  while ((1 << bits) < p->h) ++bits;
  for (int i = 0; i < p->h; ++i) {
    int r = 0;
    for (int b = 0; b < bits; ++b)
      if (i & (1 << b)) r |= 1 << (bits - 1 - b);
    p->rev[i] = r;
  }
  return p;
}
static void plan_free(plan_t* p) {
  free(p->w); free(p->wr); free(p->rev); free(p);
}

/* in-place complex FFT of size h (sign = -1 forward, +1 inverse, unnormalised) */
static void cfft(const plan_t* p, cf* a, int sign) {
  const int h = p->h;
  for (int i = 0; i < h; ++i) {
    const int r = p->rev[i];
    if (r > i) { cf t = a[i]; a[i] = a[r]; a[r] = t; }
  }
  for (int len = 2; len <= h; len <<= 1) {
    const int half = len >> 1, step = h / len;
    for (int s = 0; s < h; s += len) {
      for (int k = 0; k < half; ++k) {
        cf w = p->w[k * step];
        if (sign > 0) w.im = -w.im;
        const cf u = a[s + k], v = a[s + k + half];
        const float tr = v.re * w.re - v.im * w.im, ti = v.re * w.im + v.im * w.re;
        a[s + k].re = u.re + tr; a[s + k].im = u.im + ti;
        a[s + k + half].re = u.re - tr; a[s + k + half].im = u.im - ti;
      }
    }
  }
}

/* unnormalised r2c: x[n] -> spec[n/2+1] */
static void rfft(const plan_t* p, const float* x, cf* spec, cf* work) {
  const int h = p->h;
  for (int i = 0; i < h; ++i) { work[i].re = x[2 * i]; work[i].im = x[2 * i + 1]; }
  cfft(p, work, -1);
  for (int k = 0; k <= h; ++k) {
    const cf zk = work[k % h], zc = work[(h - k) % h];
    const float er = 0.5f * (zk.re + zc.re), ei = 0.5f * (zk.im - zc.im);      /* even part */
    const float orr = 0.5f * (zk.im + zc.im), oi = -0.5f * (zk.re - zc.re);    /* odd part  */
    const cf w = p->wr[k];
    spec[k].re = er + orr * w.re - oi * w.im;
    spec[k].im = ei + orr * w.im + oi * w.re;
  }
}

/* unnormalised c2r (imaginary parts of DC / Nyquist ignored): spec[n/2+1] -> x[n] */
static void irfft(const plan_t* p, const cf* spec, float* x, cf* work) {
  const int h = p->h;
  for (int k = 0; k < h; ++k) {
    cf a = spec[k], b = spec[h - k];
    if (k == 0) { a.im = 0.f; b.im = 0.f; }
    const float er = a.re + b.re, ei = a.im - b.im;
    const float dr = a.re - b.re, di = a.im + b.im;
    const cf w = p->wr[k];                 /* multiply the odd part by i * conj(w) */
    const float tr = dr * w.re + di * w.im, ti = di * w.re - dr * w.im;
    work[k].re = er - ti;
    work[k].im = ei + tr;
  }
  cfft(p, work, +1);
  for (int i = 0; i < h; ++i) { x[2 * i] = work[i].re; x[2 * i + 1] = work[i].im; }
}

static const float kPi = 3.14159265358979323846f;

typedef struct {
  float* data;   /* [P][N] */
  cf* fft;       /* [P][F] */
  float* amp;    /* [P][F] */
  float* phase;  /* [P][F] */
} slot_t;

static void slot_alloc(slot_t* s, int64_t P, int N) {
  const int F = N / 2 + 1;
  s->data = (float*)malloc(sizeof(float) * (size_t)P * N);
  s->fft = (cf*)calloc((size_t)P * F, sizeof(cf));
  s->amp = (float*)calloc((size_t)P * F, sizeof(float));
  s->phase = (float*)calloc((size_t)P * F, sizeof(float));
}
static void slot_free(slot_t* s) { free(s->data); free(s->fft); free(s->amp); free(s->phase); }
/* `input.clone()`: a single-threaded deep copy, as in the reference */
static void slot_clone(slot_t* dst, const slot_t* src, int64_t P, int N) {
  const int F = N / 2 + 1;
  memcpy(dst->data, src->data, sizeof(float) * (size_t)P * N);
  memcpy(dst->fft, src->fft, sizeof(cf) * (size_t)P * F);
  memcpy(dst->amp, src->amp, sizeof(float) * (size_t)P * F);
  memcpy(dst->phase, src->phase, sizeof(float) * (size_t)P * F);
}


typedef struct { const plan_t* plan; slot_t* cur; const float* window; int cols; int N; } stage_ctx;
typedef struct { const float* data; float* img; int cols; int N; } img_ctx;

/* one rayon task of `fft` (math_tools.rs:333-392): all traces of row r */
static void fft_row(int r, void* vctx) {
  stage_ctx* c = (stage_ctx*)vctx;
  const int N = c->N, F = N / 2 + 1, cols = c->cols;
  cf* work = (cf*)malloc(sizeof(cf) * (size_t)(N / 2 + 1));
  float* ph = (float*)malloc(sizeof(float) * (size_t)F);
  for (int col = 0; col < cols; ++col) {
    const int64_t p = (int64_t)r * cols + col;
    float* x = c->cur->data + p * N;
    if (c->window)
      for (int i = 0; i < N; ++i) x[i] *= c->window[i];
    cf* sp = c->cur->fft + p * F;
    rfft(c->plan, x, sp, work);
    for (int k = 0; k < F; ++k) {
      c->cur->amp[p * F + k] = hypotf(sp[k].re, sp[k].im);
      ph[k] = atan2f(sp[k].im, sp[k].re);
    }
    /* numpy_unwrap, period 2 pi (math_tools.rs:211-240) */
    float prev = ph[0], acc = ph[0];
    c->cur->phase[p * F] = acc;
    for (int k = 1; k < F; ++k) {
      float d = ph[k] - prev;
      if (d > kPi) d -= 2.0f * kPi;
      else if (d < -kPi) d += 2.0f * kPi;
      acc += d;
      prev = ph[k];
      c->cur->phase[p * F + k] = acc;
    }
  }
  free(work);
  free(ph);
}

/* one rayon task of `ifft` (math_tools.rs:546-567) */
static void ifft_row(int r, void* vctx) {
  stage_ctx* c = (stage_ctx*)vctx;
  const int N = c->N, F = N / 2 + 1, cols = c->cols;
  cf* work = (cf*)malloc(sizeof(cf) * (size_t)(N / 2 + 1));
  for (int col = 0; col < cols; ++col) {
    const int64_t p = (int64_t)r * cols + col;
    float* x = c->cur->data + p * N;
    irfft(c->plan, c->cur->fft + p * F, x, work);
    const float inv = (float)N;
    for (int i = 0; i < N; ++i) x[i] = x[i] / inv;
  }
  free(work);
}

static void img_row(int r, void* vctx) {
  img_ctx* c = (img_ctx*)vctx;
  for (int col = 0; col < c->cols; ++col) {
    const int64_t p = (int64_t)r * c->cols + col;
    const float* x = c->data + p * c->N;
    float s = 0.f;
    for (int i = 0; i < c->N; ++i) s += x[i] * x[i];
    c->img[p] = s;
  }
}

/* Runs slots 1..7 of the default chain.  Multiplier vectors are the pixel-independent factors the
 * reference recomputes per trace (taper, gates, window, band); they are passed in so that this twin
 * and the numpy oracle share one definition (the per-trace recomputation cost of ~2 cosf per tapered
 * sample is small next to the FFTs and is not modelled).
 * rows x cols pixels; outputs: data7 [P][N], img [P], and optionally fft5 / amp5 / phase4.
 * Returns 0. */
int thzc_default_chain(const float* data0, int rows, int cols, int N, const float* tilt, const float* gate_before,
                       const float* window, const float* band, const float* gate_after, float* data7, float* img,
                       float* fft5_out, float* amp5_out, float* phase4_out, int threads) {
  const int64_t P = (int64_t)rows * cols;
  const int F = N / 2 + 1;
  g_threads = threads > 0 ? threads : (int)sysconf(_SC_NPROCESSORS_ONLN);
  plan_t* plan = plan_new(N);
  slot_t cur, nxt;
  slot_alloc(&cur, P, N);
  slot_alloc(&nxt, P, N);
  memcpy(cur.data, data0, sizeof(float) * (size_t)P * N);
#define NEXT_STAGE() do { slot_clone(&nxt, &cur, P, N); slot_t t_ = cur; cur = nxt; nxt = t_; } while (0)
  /* slot 1: scaling (clone) */
  NEXT_STAGE();
  /* slot 2: tilt taper, serial pixel loop */
  NEXT_STAGE();
  if (tilt)
    for (int64_t p = 0; p < P; ++p)
      for (int i = 0; i < N; ++i) cur.data[p * N + i] *= tilt[i];
  /* slot 3: gate before FFT, serial pixel loop */
  NEXT_STAGE();
  if (gate_before)
    for (int64_t p = 0; p < P; ++p)
      for (int i = 0; i < N; ++i) cur.data[p * N + i] *= gate_before[i];
  /* slot 4: fft, parallel over axis 0 */
  NEXT_STAGE();
  {
    stage_ctx sc = {plan, &cur, window, cols, N};
    parallel_rows(fft_row, &sc, rows);
  }
  if (phase4_out) memcpy(phase4_out, cur.phase, sizeof(float) * (size_t)P * F);
  /* slot 5: FD band-pass -- the reference's parallel loop holds two mutexes for its whole body, i.e.
   * it is serial; followed by the serial zero-pad copy (modelled by the clone + in-place multiply) */
  NEXT_STAGE();
  if (band)
    for (int64_t p = 0; p < P; ++p)
      for (int k = 0; k < F; ++k) {
        cur.fft[p * F + k].re *= band[k];
        cur.fft[p * F + k].im *= band[k];
        cur.amp[p * F + k] *= band[k];
      }
  if (fft5_out) memcpy(fft5_out, cur.fft, sizeof(cf) * (size_t)P * F);
  if (amp5_out) memcpy(amp5_out, cur.amp, sizeof(float) * (size_t)P * F);
  /* slot 6: ifft -- pixel means (serial, 3 passes over the spectral cubes), then parallel over axis 0 */
  NEXT_STAGE();
  {
    double* acc = (double*)calloc((size_t)4 * F, sizeof(double));
    for (int64_t p = 0; p < P; ++p)
      for (int k = 0; k < F; ++k) {
        acc[k] += cur.fft[p * F + k].re;
        acc[F + k] += cur.fft[p * F + k].im;
        acc[2 * F + k] += cur.amp[p * F + k];
        acc[3 * F + k] += cur.phase[p * F + k];
      }
    free(acc);
  }
  {
    stage_ctx sc = {plan, &cur, NULL, cols, N};
    parallel_rows(ifft_row, &sc, rows);
  }
  /* slot 7: gate after iFFT, serial pixel loop */
  NEXT_STAGE();
  if (gate_after)
    for (int64_t p = 0; p < P; ++p)
      for (int i = 0; i < N; ++i) cur.data[p * N + i] *= gate_after[i];
  /* slot 8 (deconvolution inactive): clone */
  NEXT_STAGE();
  /* intensity image, parallel over rows */
  {
    img_ctx ic = {cur.data, img, cols, N};
    parallel_rows(img_row, &ic, rows);
  }
  memcpy(data7, cur.data, sizeof(float) * (size_t)P * N);
  slot_free(&cur);
  slot_free(&nxt);
  plan_free(plan);
  return 0;
}

int thzc_num_threads(void) { return (int)sysconf(_SC_NPROCESSORS_ONLN); }
