/*
 * rr_oracle.c — CPU restatement of rustradio's filtering hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under rustradio_b200/ (the product) may
 * import, link or call this file.  It is used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * legs, and there only as the checker / the CPU baseline.
 *
 * The reference (ThomasHabets/rustradio v0.18.2) is Rust and there is no Rust
 * toolchain in this image, so the reference itself cannot be compiled into
 * oracle/_ref.  Every function below follows the reference lines it cites
 * (paths relative to /root/reference).  Pinning status:
 *   - FIR, resampler, demod, tap design, count rules: pinned by the
 *     reference's own known-answer tests (tests/test_oracle_kat.py).
 *   - FFT filter numerics: the FFT is rustfft 6.4.1 (Cargo.lock:2299-2300,
 *     not vendored) -> bit pattern "parity unpinned"; pinned only through
 *     the reference's property tests (filter_a_signal, tag_propagation) and
 *     against an f64 direct convolution.
 *   - QuadratureDemod default build uses fast-math 0.1.1 atan2
 *     (Cargo.lock:815-816, not vendored); this file restates the libm
 *     branch (src/quadrature_demod.rs:96-108) -> the fast approximation is
 *     "parity unpinned" and deliberately not reproduced.
 *
 * Build: see oracle/Makefile.  Must be compiled with -ffp-contract=off
 * (rustc never contracts a*b+c into an FMA).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float re, im; } c32;
typedef struct { double re, im; } c64;

/* num-complex 0.4.6 Mul: (a+bi)(c+di) = (ac-bd) + (ad+bc)i, each op rounded. */
static inline c32 c32_mul(c32 a, c32 b) {
    c32 r;
    r.re = a.re * b.re - a.im * b.im;
    r.im = a.re * b.im + a.im * b.re;
    return r;
}
static inline c32 c32_add(c32 a, c32 b) {
    c32 r = { a.re + b.re, a.im + b.im };
    return r;
}

/* ------------------------------------------------------------------ FIR -- */

/*
 * FirFilter output count for a whole stream of n inputs:
 * src/fir.rs:496-525 applied until WaitForStream; see SURVEY Appendix A.
 */
int64_t orc_fir_out_count(int64_t n, int64_t ntaps, int64_t deci) {
    if (n < ntaps + deci - 1) return 0;
    return (n - ntaps + 1) / deci;
}

/*
 * Fir::new + Fir::filter + Fir::filter_n_inplace, T = Complex.
 * src/fir.rs:156-162 (taps stored reversed), :166-177 (fold, ascending j,
 * acc + tap*input), :192-197 (stride deci).  `taps` is given in the caller's
 * (un-reversed) order, exactly as passed to Fir::new.
 */
void orc_fir_c32(const c32* x, const c32* taps, int64_t ntaps, int64_t deci,
                 c32* out, int64_t out_n) {
    c32* rev = (c32*)malloc(sizeof(c32) * (size_t)ntaps);
    for (int64_t j = 0; j < ntaps; j++) rev[j] = taps[ntaps - 1 - j];
    for (int64_t i = 0; i < out_n; i++) {
        const c32* in = x + i * deci;
        c32 acc = { 0.0f, 0.0f };
        for (int64_t j = 0; j < ntaps; j++) acc = c32_add(acc, c32_mul(rev[j], in[j]));
        out[i] = acc;
    }
    free(rev);
}

/* Same, T = Float (src/fir.rs:166-177 generic over T). */
void orc_fir_f32(const float* x, const float* taps, int64_t ntaps, int64_t deci,
                 float* out, int64_t out_n) {
    float* rev = (float*)malloc(sizeof(float) * (size_t)ntaps);
    for (int64_t j = 0; j < ntaps; j++) rev[j] = taps[ntaps - 1 - j];
    for (int64_t i = 0; i < out_n; i++) {
        const float* in = x + i * deci;
        float acc = 0.0f;
        for (int64_t j = 0; j < ntaps; j++) acc = acc + rev[j] * in[j];
        out[i] = acc;
    }
    free(rev);
}

/* f64 ground truth of the same sums (inputs are the same f32 values). */
void orc_fir_c32_f64(const c32* x, const c32* taps, int64_t ntaps, int64_t deci,
                     c64* out, int64_t out_n) {
    for (int64_t i = 0; i < out_n; i++) {
        const c32* in = x + i * deci;
        double re = 0.0, im = 0.0;
        for (int64_t j = 0; j < ntaps; j++) {
            c32 t = taps[ntaps - 1 - j];
            re += (double)t.re * in[j].re - (double)t.im * in[j].im;
            im += (double)t.re * in[j].im + (double)t.im * in[j].re;
        }
        out[i].re = re; out[i].im = im;
    }
}
void orc_fir_f32_f64(const float* x, const float* taps, int64_t ntaps, int64_t deci,
                     double* out, int64_t out_n) {
    for (int64_t i = 0; i < out_n; i++) {
        const float* in = x + i * deci;
        double acc = 0.0;
        for (int64_t j = 0; j < ntaps; j++) acc += (double)taps[ntaps - 1 - j] * in[j];
        out[i] = acc;
    }
}

/*
 * ComplexFrequencyTranslator::new_translator, src/fir.rs:427-462.
 * Rotates `taps` in place by the f32 recurrence and returns the initial
 * output phase and the per-output step.  freq == 0 -> returns 0 (no
 * translator, :438-440) and leaves taps untouched.
 */
int orc_fir_new_translator(c32* taps, int64_t ntaps, float samp_rate, float freq,
                           int64_t deci, c32* phase0, c32* step) {
    if (freq == 0.0f) return 0;
    double input_step = 2.0 * M_PI * (double)freq / (double)samp_rate;
    c32 tap_step = { (float)cos(input_step), (float)sin(input_step) };
    c32 phase = { 1.0f, 0.0f };
    for (int64_t k = 0; k < ntaps; k++) {
        taps[k] = c32_mul(taps[k], phase);   /* *tap *= phase  */
        phase = c32_mul(phase, tap_step);    /* phase *= tap_step */
    }
    double first_output_phase = -input_step * (double)(ntaps - 1);
    double output_step = -input_step * (double)deci;
    phase0->re = (float)cos(first_output_phase);
    phase0->im = (float)sin(first_output_phase);
    step->re = (float)cos(output_step);
    step->im = (float)sin(output_step);
    return 1;
}

/* translate_output, src/fir.rs:464-473.  `phase` is carried across calls. */
void orc_fir_translate_output(c32* out, int64_t n, c32* phase, const c32* step) {
    for (int64_t i = 0; i < n; i++) {
        out[i] = c32_mul(out[i], *phase);
        *phase = c32_mul(*phase, *step);
    }
}

/* ------------------------------------------------------ tap design ------ */

/* src/window.rs:98-112 (Hamming, f32 arithmetic, PI as f32). */
static void win_hamming(float* w, int64_t ntaps, float a0) {
    const float PI = (float)M_PI;
    if (ntaps == 1) { w[0] = 1.0f; return; }
    float a1 = 1.0f - a0;
    float m = (float)(ntaps - 1);
    for (int64_t n = 0; n < ntaps; n++) w[n] = a0 - a1 * cosf(2.0f * PI * (float)n / m);
}
/* src/window.rs:117-153 (Blackman, A=0.16; divides by m, not m-1). */
static void win_blackman(float* w, int64_t m_) {
    const float PI = (float)M_PI;
    const float A = 0.16f;
    if (m_ == 1) { w[0] = 1.0f; return; }
    float m = (float)m_;
    float a0 = (1.0f - A) / 2.0f, a1 = 0.5f, a2 = A / 2.0f;
    for (int64_t i = 0; i < m_; i++) {
        float n = (float)i;
        float t1 = 2.0f * PI * n / m, t2 = 4.0f * PI * n / m;
        w[i] = a0 - a1 * cosf(t1) + a2 * cosf(t2);
    }
}
/* src/window.rs:158-185 (Blackman-Harris). */
static void win_blackman_harris(float* w, int64_t m_) {
    const float PI = (float)M_PI;
    const float A0 = 0.35875f, A1 = 0.48829f, A2 = 0.14128f, A3 = 0.01168f;
    if (m_ == 1) { w[0] = 1.0f; return; }
    float m = (float)m_;
    for (int64_t i = 0; i < m_; i++) {
        float n = (float)i;
        float t1 = 2.0f * PI * n / m, t2 = 4.0f * PI * n / m, t3 = 6.0f * PI * n / m;
        w[i] = A0 - A1 * cosf(t1) + A2 * cosf(t2) - A3 * cosf(t3);
    }
}

/* window_type: 0 Hamming (a0 = 25/46), 1 Blackman, 2 BlackmanHarris,
 * 3 HammingParm(parm).  src/window.rs:36-37,63-86. */
int orc_make_window(int window_type, float parm, int64_t ntaps, float* w) {
    if (ntaps <= 0) return 0;
    switch (window_type) {
    case 0: win_hamming(w, ntaps, 25.0f / 46.0f); return 0;
    case 1: win_blackman(w, ntaps); return 0;
    case 2: win_blackman_harris(w, ntaps); return 0;
    case 3: win_hamming(w, ntaps, parm); return 0;
    default: return -1;
    }
}
static float win_max_attenuation(int window_type) {
    switch (window_type) { case 1: return 74.0f; case 2: return 92.0f; default: return 53.0f; }
}

/* compute_ntaps, src/fir.rs:606-610 (always odd). */
int64_t orc_compute_ntaps(float samp_rate, float twidth, int window_type) {
    float a = win_max_attenuation(window_type);
    int64_t t = (int64_t)(a * samp_rate / (22.0f * twidth));
    return (t & 1) == 0 ? t + 1 : t;
}

/*
 * low_pass body, src/fir.rs:631-655, generalised to an explicit ntaps (the
 * reference always derives ntaps from compute_ntaps; BASELINE configs need
 * exact even/odd lengths).  All arithmetic in f32 like the
 * reference.  For odd ntaps from compute_ntaps this is the reference formula.
 */
void orc_low_pass_n(float samp_rate, float cutoff, int window_type, float parm,
                    int64_t ntaps, float* taps) {
    const float pi = (float)M_PI;
    float* window = (float*)malloc(sizeof(float) * (size_t)ntaps);
    orc_make_window(window_type, parm, ntaps, window);
    int64_t m = (ntaps - 1) / 2;
    float fwt0 = 2.0f * pi * cutoff / samp_rate;
    for (int64_t nm = 0; nm < ntaps; nm++) {
        int64_t n = nm - m;
        float nf = (float)n;
        if (n == 0) taps[nm] = fwt0 / pi * window[nm];
        else taps[nm] = (sinf(nf * fwt0) / (nf * pi)) * window[nm];
    }
    float fmax = taps[m];
    for (int64_t n = 1; n <= m; n++) fmax += 2.0f * taps[n + m];
    float gain = 1.0f / fmax;
    for (int64_t i = 0; i < ntaps; i++) taps[i] = taps[i] * gain;
    free(window);
}

/* --------------------------------------------------------- FFT filter ---- */

/* calc_fft_size, src/fft_filter.rs:36-42. */
int64_t orc_calc_fft_size(int64_t from) {
    int64_t n = 1;
    while (n < from) n <<= 1;
    return 2 * n;
}
/* FftFilter whole-stream output count: blocks of nsamples = fft_size - ntaps,
 * trailing partial block never flushed (src/fft_filter.rs:315-327). */
int64_t orc_fftfilt_out_count(int64_t n, int64_t ntaps) {
    int64_t s = orc_calc_fft_size(ntaps) - ntaps;
    return (n / s) * s;
}

/*
 * f32 complex FFT used in place of rustfft (un-vendored): Stockham autosort,
 * radix-4 passes plus one radix-2 pass when log2(n) is odd; twiddles computed
 * in f64 and rounded to f32 once.  Unnormalised, like rustfft.
 * Split (SoA) re/im arrays so the inner loops vectorise.
 */
typedef struct {
    int64_t n;
    float* twr; float* twi;      /* n entries: e^{-2 pi i k / n} */
    float* ar; float* ai; float* br; float* bi;   /* work buffers */
} orc_fft_plan;

static orc_fft_plan* fft_plan_new(int64_t n) {
    orc_fft_plan* p = (orc_fft_plan*)malloc(sizeof(*p));
    p->n = n;
    size_t sz = sizeof(float) * (size_t)n;
    p->twr = (float*)aligned_alloc(64, sz > 64 ? sz : 64); p->twi = (float*)aligned_alloc(64, sz > 64 ? sz : 64);
    p->ar = (float*)aligned_alloc(64, sz > 64 ? sz : 64);  p->ai = (float*)aligned_alloc(64, sz > 64 ? sz : 64);
    p->br = (float*)aligned_alloc(64, sz > 64 ? sz : 64);  p->bi = (float*)aligned_alloc(64, sz > 64 ? sz : 64);
    for (int64_t k = 0; k < n; k++) {
        double a = -2.0 * M_PI * (double)k / (double)n;
        p->twr[k] = (float)cos(a); p->twi[k] = (float)sin(a);
    }
    return p;
}
static void fft_plan_free(orc_fft_plan* p) {
    free(p->twr); free(p->twi); free(p->ar); free(p->ai); free(p->br); free(p->bi); free(p);
}

/* One Stockham radix-2 pass: length-n sub-transforms at stride s. */
static void stockham_r2(const orc_fft_plan* pl, int64_t n, int64_t s, int sign,
                        const float* xr, const float* xi, float* yr, float* yi) {
    int64_t m = n / 2, tstep = pl->n / n;
    for (int64_t p = 0; p < m; p++) {
        float wr = pl->twr[p * tstep], wi = sign * pl->twi[p * tstep];
        for (int64_t q = 0; q < s; q++) {
            float a_r = xr[q + s * p], a_i = xi[q + s * p];
            float b_r = xr[q + s * (p + m)], b_i = xi[q + s * (p + m)];
            yr[q + s * (2 * p)] = a_r + b_r; yi[q + s * (2 * p)] = a_i + b_i;
            float dr = a_r - b_r, di = a_i - b_i;
            yr[q + s * (2 * p + 1)] = dr * wr - di * wi;
            yi[q + s * (2 * p + 1)] = dr * wi + di * wr;
        }
    }
}
/* One Stockham radix-4 pass. sign = +1 forward (e^{-i..}), -1 inverse. */
static void stockham_r4(const orc_fft_plan* pl, int64_t n, int64_t s, int sign,
                        const float* xr, const float* xi, float* yr, float* yi) {
    int64_t n1 = n / 4, tstep = pl->n / n;
    for (int64_t p = 0; p < n1; p++) {
        float w1r = pl->twr[p * tstep],     w1i = sign * pl->twi[p * tstep];
        float w2r = pl->twr[2 * p * tstep], w2i = sign * pl->twi[2 * p * tstep];
        float w3r = pl->twr[3 * p * tstep], w3i = sign * pl->twi[3 * p * tstep];
        for (int64_t q = 0; q < s; q++) {
            float ar = xr[q + s * p],            ai = xi[q + s * p];
            float br = xr[q + s * (p + n1)],     bi = xi[q + s * (p + n1)];
            float cr = xr[q + s * (p + 2 * n1)], ci = xi[q + s * (p + 2 * n1)];
            float dr = xr[q + s * (p + 3 * n1)], di = xi[q + s * (p + 3 * n1)];
            float apc_r = ar + cr, apc_i = ai + ci, amc_r = ar - cr, amc_i = ai - ci;
            float bpd_r = br + dr, bpd_i = bi + di;
            /* -i*(b-d) for forward, +i*(b-d) for inverse */
            float bmd_r = br - dr, bmd_i = bi - di;
            float jr = sign * bmd_i, ji = -sign * bmd_r;   /* (-i*sign)*(b-d) */
            yr[q + s * (4 * p)] = apc_r + bpd_r; yi[q + s * (4 * p)] = apc_i + bpd_i;
            float t1r = amc_r + jr, t1i = amc_i + ji;
            float t2r = apc_r - bpd_r, t2i = apc_i - bpd_i;
            float t3r = amc_r - jr, t3i = amc_i - ji;
            yr[q + s * (4 * p + 1)] = t1r * w1r - t1i * w1i; yi[q + s * (4 * p + 1)] = t1r * w1i + t1i * w1r;
            yr[q + s * (4 * p + 2)] = t2r * w2r - t2i * w2i; yi[q + s * (4 * p + 2)] = t2r * w2i + t2i * w2r;
            yr[q + s * (4 * p + 3)] = t3r * w3r - t3i * w3i; yi[q + s * (4 * p + 3)] = t3r * w3i + t3i * w3r;
        }
    }
}
/* In-place (on interleaved c32 buf) unnormalised FFT. sign=+1 fwd, -1 inv. */
static void fft_run(orc_fft_plan* pl, c32* buf, int sign) {
    int64_t N = pl->n;
    for (int64_t i = 0; i < N; i++) { pl->ar[i] = buf[i].re; pl->ai[i] = buf[i].im; }
    float *xr = pl->ar, *xi = pl->ai, *yr = pl->br, *yi = pl->bi;
    int64_t n = N, s = 1;
    while (n > 1) {
        if (n % 4 == 0) { stockham_r4(pl, n, s, sign, xr, xi, yr, yi); n /= 4; s *= 4; }
        else            { stockham_r2(pl, n, s, sign, xr, xi, yr, yi); n /= 2; s *= 2; }
        float* t;
        t = xr; xr = yr; yr = t;
        t = xi; xi = yi; yi = t;
    }
    for (int64_t i = 0; i < N; i++) { buf[i].re = xr[i]; buf[i].im = xi[i]; }
}

/* Stand-alone transform for tests (out-of-place convenience). */
void orc_fft_c32(c32* buf, int64_t n, int inverse) {
    orc_fft_plan* p = fft_plan_new(n);
    fft_run(p, buf, inverse ? -1 : +1);
    fft_plan_free(p);
}

/*
 * FftFilter state + work loop, src/fft_filter.rs:144-176 (engine),
 * :259-278 (sizes, zero tail), :331-348 (overlap-ADD block).
 */
typedef struct {
    int64_t ntaps, fft_size, nsamples;
    c32* taps_fft;     /* FFT(zero-padded taps) * (1/fft_size), :153-162 */
    c32* tail;         /* ntaps entries, starts zero, :270 */
    c32* buf;          /* fft_size */
    orc_fft_plan* plan;
} orc_fftfilt;

orc_fftfilt* orc_fftfilt_new(const c32* taps, int64_t ntaps) {
    orc_fftfilt* f = (orc_fftfilt*)calloc(1, sizeof(*f));
    f->ntaps = ntaps;
    f->fft_size = orc_calc_fft_size(ntaps);
    f->nsamples = f->fft_size - ntaps;
    f->plan = fft_plan_new(f->fft_size);
    f->taps_fft = (c32*)calloc((size_t)f->fft_size, sizeof(c32));
    memcpy(f->taps_fft, taps, sizeof(c32) * (size_t)ntaps);
    fft_run(f->plan, f->taps_fft, +1);
    float scale = 1.0f / (float)f->fft_size;
    for (int64_t i = 0; i < f->fft_size; i++) {   /* *s *= f (Complex *= Float) */
        f->taps_fft[i].re *= scale; f->taps_fft[i].im *= scale;
    }
    f->tail = (c32*)calloc((size_t)ntaps, sizeof(c32));
    f->buf = (c32*)calloc((size_t)f->fft_size, sizeof(c32));
    return f;
}
void orc_fftfilt_free(orc_fftfilt* f) {
    fft_plan_free(f->plan); free(f->taps_fft); free(f->tail); free(f->buf); free(f);
}
int64_t orc_fftfilt_nsamples(const orc_fftfilt* f) { return f->nsamples; }
int64_t orc_fftfilt_fft_size(const orc_fftfilt* f) { return f->fft_size; }

/* Process `nblocks` whole blocks (x has nblocks*nsamples samples). */
void orc_fftfilt_run(orc_fftfilt* f, const c32* x, int64_t nblocks, c32* out) {
    const int64_t S = f->nsamples, F = f->fft_size, T = f->ntaps;
    for (int64_t b = 0; b < nblocks; b++) {
        memcpy(f->buf, x + b * S, sizeof(c32) * (size_t)S);
        memset(f->buf + S, 0, sizeof(c32) * (size_t)(F - S));          /* resize(fft_size, 0) :332 */
        fft_run(f->plan, f->buf, +1);                                  /* :173 */
        for (int64_t i = 0; i < F; i++) f->buf[i] = c32_mul(f->buf[i], f->taps_fft[i]);  /* sum_vec :281-287 */
        fft_run(f->plan, f->buf, -1);                                  /* :175 */
        for (int64_t i = 0; i < T; i++) f->buf[i] = c32_add(f->buf[i], f->tail[i]);      /* :336-338 */
        memcpy(out + b * S, f->buf, sizeof(c32) * (size_t)S);          /* :342 */
        for (int64_t i = 0; i < T; i++) f->tail[i] = f->buf[S + i];    /* :346-348 */
    }
}

/* f64 truth: y[n] = sum_k h[k] x[n-k], x[n<0] = 0.
 * Direct O(n_out * ntaps); use only at sizes that finish in seconds. */
void orc_conv_full_c32_f64(const c32* x, int64_t n_out, const c32* h, int64_t ntaps, c64* out) {
    for (int64_t n = 0; n < n_out; n++) {
        double re = 0.0, im = 0.0;
        int64_t kmax = n < ntaps - 1 ? n : ntaps - 1;
        for (int64_t k = 0; k <= kmax; k++) {
            c32 t = h[k], v = x[n - k];
            re += (double)t.re * v.re - (double)t.im * v.im;
            im += (double)t.re * v.im + (double)t.im * v.re;
        }
        out[n].re = re; out[n].im = im;
    }
}

/* ---------------------------------------------------- RationalResampler -- */

static int64_t gcd_i64(int64_t a, int64_t b) {   /* src/rational_resampler.rs:10-17 */
    while (b != 0) { int64_t t = b; b = a % b; a = t; }
    return a;
}

typedef struct {
    int64_t deci, interp, counter;
    int has_pending;
    uint8_t pending[16];
    int64_t elem;
} orc_resampler;

/* RationalResampler::new, src/rational_resampler.rs:125-151.  NULL on 0. */
orc_resampler* orc_resampler_new(int64_t elem_size, int64_t interp, int64_t deci) {
    if (deci == 0 || interp == 0 || elem_size <= 0 || elem_size > 16) return NULL;
    int64_t g = gcd_i64(deci, interp);
    orc_resampler* r = (orc_resampler*)calloc(1, sizeof(*r));
    r->deci = deci / g; r->interp = interp / g; r->counter = 0; r->has_pending = 0;
    r->elem = elem_size;
    return r;
}
void orc_resampler_free(orc_resampler* r) { free(r); }
int orc_resampler_has_pending(const orc_resampler* r) { return r->has_pending; }
int64_t orc_resampler_counter(const orc_resampler* r) { return r->counter; }
/* Test hook: preset the struct field `counter` (src/rational_resampler.rs:101-105) so that the work() loop
 * below — unchanged — starts mid-stream, the way a time-segment shard does. */
void orc_resampler_set_counter(orc_resampler* r, int64_t counter) { r->counter = counter; r->has_pending = 0; }

/*
 * One RationalResampler::work() call, src/rational_resampler.rs:155-206, on
 * an input window of n_in samples and an output window of out_cap samples.
 * Returns 0 = WaitForStream(dst,1), 1 = WaitForStream(src,1).
 */
int orc_resampler_work(orc_resampler* r, const void* in, int64_t n_in,
                       void* out, int64_t out_cap, int64_t* consumed, int64_t* produced) {
    const uint8_t* ip = (const uint8_t*)in;
    uint8_t* op = (uint8_t*)out;
    const int64_t E = r->elem;
    int64_t opos = 0;
    *consumed = 0; *produced = 0;
    if (out_cap == 0) return 0;                                  /* :157-159 */
    if (r->has_pending) {                                        /* :161-173 */
        while (r->counter > 0) {
            memcpy(op + opos * E, r->pending, (size_t)E);
            r->counter -= r->deci;
            opos++;
            if (opos == out_cap) { *produced = opos; return 0; }
        }
        r->has_pending = 0;
    }
    if (n_in == 0) { *produced = opos; return 1; }               /* :175-179 */
    int64_t taken = 0; int out_full = 0;
    for (int64_t s = 0; s < n_in && !out_full; s++) {            /* :183-198 */
        taken++;
        r->counter += r->interp;
        while (r->counter > 0) {
            memcpy(op + opos * E, ip + s * E, (size_t)E);
            r->counter -= r->deci;
            opos++;
            if (opos == out_cap) {
                out_full = 1;
                if (r->counter > 0) { r->has_pending = 1; memcpy(r->pending, ip + s * E, (size_t)E); }
                break;
            }
        }
    }
    *consumed = taken; *produced = opos;
    return out_full ? 0 : 1;
}

/* Whole-stream count: ceil(N*I'/D'). */
int64_t orc_resample_out_count(int64_t n, int64_t interp, int64_t deci) {
    int64_t g = gcd_i64(deci, interp);
    interp /= g; deci /= g;
    __int128 num = (__int128)n * interp;
    return (int64_t)((num + deci - 1) / deci);
}

/* ------------------------------------------------------ QuadratureDemod -- */

/*
 * src/quadrature_demod.rs:71-73 (tmp = conj(x[t]) * x[t+1]) and :106-108
 * (libm branch: gain * im.atan2(re)).  n_in samples -> n_in-1 outputs.
 */
void orc_quad_demod(const c32* x, int64_t n_in, float gain, float* out) {
    for (int64_t t = 0; t + 1 < n_in; t++) {
        c32 a = { x[t].re, -x[t].im };          /* conj */
        c32 p = c32_mul(a, x[t + 1]);
        out[t] = gain * atan2f(p.im, p.re);
    }
}
/* f64 truth of the angle (before gain), for the <=1e-4 rad bar. */
void orc_quad_demod_f64(const c32* x, int64_t n_in, double gain, double* out) {
    for (int64_t t = 0; t + 1 < n_in; t++) {
        double ar = x[t].re, ai = -(double)x[t].im, br = x[t + 1].re, bi = x[t + 1].im;
        out[t] = gain * atan2(ar * bi + ai * br, ar * br - ai * bi);
    }
}

/* --------------------------------------------------------- RtlSdrDecode -- */

/*
 * src/rtlsdr_decode.rs:35-43: pairs of bytes (I, Q) ->
 * Complex((I - 127.0) * 0.008, (Q - 127.0) * 0.008), all in f32, subtraction
 * rounded before the multiplication.  n_bytes & !1 bytes are used (:23).
 */
void orc_rtlsdr_decode(const uint8_t* in, int64_t n_bytes, c32* out) {
    for (int64_t i = 0; i + 1 < n_bytes; i += 2) {
        float a = (float)in[i], b = (float)in[i + 1];
        out[i / 2].re = (a - 127.0f) * 0.008f;
        out[i / 2].im = (b - 127.0f) * 0.008f;
    }
}

/* --------------------------------------------------------- RtlSdrEncode -- */

/*
 * src/rtlsdr_encode.rs:22-26,45-46: Complex -> (u8 I, u8 Q),
 * ((sample / 0.008) + 127.0).round().clamp(0.0, 255.0) as u8 in f32; f32::round is
 * half away from zero (roundf); clamp keeps NaN, the saturating `as u8` maps it to 0.
 */
static uint8_t orc_encode_sample(float s) {
    volatile float q = s / 0.008f;                  /* volatile: keep the two roundings apart (no contraction) */
    float v = roundf(q + 127.0f);
    if (v != v) return 0;
    if (v < 0.0f) v = 0.0f;
    if (v > 255.0f) v = 255.0f;
    return (uint8_t)v;
}
void orc_rtlsdr_encode(const c32* in, int64_t n, uint8_t* out) {
    for (int64_t i = 0; i < n; ++i) {
        out[2 * i] = orc_encode_sample(in[i].re);
        out[2 * i + 1] = orc_encode_sample(in[i].im);
    }
}

/* -------------------------------------------------------------- Hilbert -- */

/* fir::hilbert, src/fir.rs:660-680 (window from orc_make_window). */
int orc_hilbert_taps(const float* window, int64_t ntaps, float* taps) {
    if (ntaps < 2) return -1;                       /* :661-662 asserts */
    int64_t mid = (ntaps - 1) / 2;
    float gain = 0.0f;
    for (int64_t i = 0; i < ntaps; i++) taps[i] = 0.0f;
    for (int64_t i = 1; i <= mid; i++) {
        if (i & 1) {
            float x = 1.0f / (float)i;
            taps[mid + i] = x * window[mid + i];
            taps[mid - i] = -x * window[mid - i];
            gain = taps[mid + i] - gain;
        } else {
            taps[mid + i] = 0.0f;
            taps[mid - i] = 0.0f;
        }
    }
    gain = 1.0f / (2.0f * fabsf(gain));
    for (int64_t i = 0; i < ntaps; i++) taps[i] = gain * taps[i];
    return 0;
}

/*
 * One Hilbert::work() pass, src/hilbert.rs:86-125: iv = history ++ input[..n];
 * out[i] = Complex(iv[i + ntaps/2], filter(iv[i..i+ntaps])) for i < n (n = len - ntaps with
 * history.len() == ntaps); history <- iv[n..len].  `history` (ntaps floats, zeros at start, :52)
 * is updated in place.  filter = Fir::filter (the scalar fallback of filter_float, src/fir.rs:145;
 * its AVX / simd branches sum in a different order and are build dependent).
 * f64_out (optional): the same sums accumulated in f64.
 */
void orc_hilbert_work(float* history, const float* taps, int64_t ntaps, const float* in, int64_t n,
                      c32* out, c64* f64_out) {
    int64_t len = ntaps + n;
    float* iv = (float*)malloc(sizeof(float) * (size_t)len);
    float* rev = (float*)malloc(sizeof(float) * (size_t)ntaps);
    memcpy(iv, history, sizeof(float) * (size_t)ntaps);
    memcpy(iv + ntaps, in, sizeof(float) * (size_t)n);
    for (int64_t j = 0; j < ntaps; j++) rev[j] = taps[ntaps - 1 - j];   /* Fir::new, src/fir.rs:156-162 */
    for (int64_t i = 0; i < n; i++) {
        if (out) {
            float acc = 0.0f;
            for (int64_t j = 0; j < ntaps; j++) acc = acc + rev[j] * iv[i + j];
            out[i].re = iv[i + ntaps / 2];
            out[i].im = acc;
        }
        if (f64_out) {
            double acc64 = 0.0;
            for (int64_t j = 0; j < ntaps; j++) acc64 += (double)rev[j] * (double)iv[i + j];
            f64_out[i].re = iv[i + ntaps / 2];
            f64_out[i].im = acc64;
        }
    }
    memcpy(history, iv + n, sizeof(float) * (size_t)ntaps);
    free(iv);
    free(rev);
}

/* ------------------------------------------------ sample-wise neighbours -- */

/* MultiplyConst::process_sync, src/multiply_const.rs:20-22 (x * val). */
void orc_multiply_const_f32(const float* x, int64_t n, float val, float* out) {
    for (int64_t i = 0; i < n; i++) out[i] = x[i] * val;
}
void orc_multiply_const_c32(const c32* x, int64_t n, float val_re, float val_im, c32* out) {
    c32 val = { val_re, val_im };
    for (int64_t i = 0; i < n; i++) out[i] = c32_mul(x[i], val);
}
/* AddConst::process_sync, src/add_const.rs:41-43 (a + val). */
void orc_add_const_f32(const float* x, int64_t n, float val, float* out) {
    for (int64_t i = 0; i < n; i++) out[i] = x[i] + val;
}
void orc_add_const_c32(const c32* x, int64_t n, float val_re, float val_im, c32* out) {
    c32 val = { val_re, val_im };
    for (int64_t i = 0; i < n; i++) out[i] = c32_add(x[i], val);
}
/* ComplexToMag2::process_sync, src/complex_to_mag2.rs:17-19: norm_sqr = re*re + im*im. */
void orc_complex_to_mag2(const c32* x, int64_t n, float* out) {
    for (int64_t i = 0; i < n; i++) out[i] = x[i].re * x[i].re + x[i].im * x[i].im;
}

/* IqBalance::with_tau alpha, src/iq_balance.rs:41-57. */
float orc_iq_balance_alpha_from_tau(uint32_t sample_rate, double tau_seconds) {
    double fs = (double)(sample_rate > 1 ? sample_rate : 1);
    double tau = (isfinite(tau_seconds) && tau_seconds > 0.0) ? tau_seconds : 0.5;
    double a = 1.0 - exp(-1.0 / (tau * fs));
    if (a < 0.0) a = 0.0;
    if (a > 1.0) a = 1.0;
    return (float)a;
}
/*
 * IqBalance::process_sync over n samples, src/iq_balance.rs:75-80:
 *   mean = mean * one_minus_alpha + x * alpha;  out = x - mean
 * (Complex * f32 scales both parts; every operation rounded to f32).  alpha is clamped and
 * one_minus_alpha = 1.0 - alpha as in with_alpha (:62-73).  `mean` is carried in place.
 */
void orc_iq_balance(const c32* x, int64_t n, float alpha, c32* mean, c32* out) {
    if (alpha < 0.0f) alpha = 0.0f;
    if (alpha > 1.0f) alpha = 1.0f;
    float oma = 1.0f - alpha;
    c32 m = *mean;
    for (int64_t i = 0; i < n; i++) {
        c32 a = { m.re * oma, m.im * oma }, b = { x[i].re * alpha, x[i].im * alpha };
        m = c32_add(a, b);
        out[i].re = x[i].re - m.re;
        out[i].im = x[i].im - m.im;
    }
    *mean = m;
}
/* The same recurrence in f64 (from the f32 alpha and 1 - alpha the block stores). */
void orc_iq_balance_f64(const c32* x, int64_t n, float alpha, c64* mean, c64* out) {
    if (alpha < 0.0f) alpha = 0.0f;
    if (alpha > 1.0f) alpha = 1.0f;
    double a = (double)alpha, oma = (double)(1.0f - alpha);
    c64 m = *mean;
    for (int64_t i = 0; i < n; i++) {
        m.re = m.re * oma + (double)x[i].re * a;
        m.im = m.im * oma + (double)x[i].im * a;
        out[i].re = (double)x[i].re - m.re;
        out[i].im = (double)x[i].im - m.im;
    }
    *mean = m;
}

/* --------------------------------------------- test-fixture restatements -- */

/* SignalSourceComplex iterator, src/signal_source.rs:39-51. `current` carried. */
void orc_signal_source_complex(float samp_rate, float freq, float amplitude,
                               double* current, c32* out, int64_t n) {
    double rad_per_sample = 2.0 * M_PI * (double)freq / (double)samp_rate;
    for (int64_t i = 0; i < n; i++) {
        *current = fmod(*current + rad_per_sample, 2.0 * M_PI);
        out[i].re = amplitude * (float)sin(*current);
        out[i].im = amplitude * (float)sin(*current - M_PI / 2.0);
    }
}

/* Deterministic synthetic input: splitmix64 counter -> U(-1,1).
 * Shared definition with the CUDA generator (rustradio_b200/csrc). */
static inline uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
void orc_synth_f32(uint64_t seed, uint64_t first_index, float* out, int64_t n) {
    for (int64_t i = 0; i < n; i++) {
        uint64_t r = splitmix64(seed ^ ((first_index + (uint64_t)i) * 0xD1342543DE82EF95ull));
        /* 24 random bits -> [0,1) -> (-1,1) */
        out[i] = (float)(r >> 40) * (1.0f / 8388608.0f) - 1.0f;
    }
}
