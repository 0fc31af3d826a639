/*
 * apt_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's shared object.  The product package never does.
 *
 * What it restates (all paths relative to /root/reference/audio_processing_tools):
 *   audio_io.py:34-72                        safe_to_float (int16 -> f32 / 32767)
 *   edge/rain_signal_processor.py:818-828    librosa.stft call (librosa 0.11.0, un-vendored:
 *                                            restated from its public semantics; PARITY UNPINNED
 *                                            at that boundary -- no reference test pins it)
 *   edge/rain_signal_processor.py:555-721    stochastic-quantile noise-PSD tracker (2 passes)
 *   edge/rain_signal_processor.py:859-888    detector dB normalisation
 *   edge/rain_frame_classifier.py:31-82      causal_stochastic_low_quantile_baseline
 *   edge/rain_frame_classifier.py:230-284    _rain_frame_decision
 *   edge/rain_frame_classifier.py:713-759    t-vs-(t-2) positive flux per mode band
 *   edge/rain_frame_classifier.py:873-998    normalisation, TD gate, NOISE/UNCERTAIN/RAIN
 *   edge/feature_extraction.py:174-538       extract_td_features_inline (default td_input_mode)
 *   edge/feature_extraction.py:542-747       extract_raw_spectral_shape_features_inline
 *   scipy.signal.sosfiltfilt / sosfilt       (scipy 1.16.3 pinned; algorithm restated)
 *   scipy.signal.peak_widths / peak_prominences
 *   numpy float32 ufunc semantics that decide bits: pairwise add.reduce, complex64 abs,
 *   AVX512 SVML log10f / log1pf (numpy/SVML, the path numpy takes on AVX512_SKX hosts,
 *   which is where the golden vectors were generated).
 *
 * The restatement is pinned against the golden vectors under tests/golden/ that were
 * produced by running the unmodified reference (oracle/make_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -pthread -shared -fPIC  (oracle/Makefile)
 * Every rounding below is deliberate; do not enable contraction or fast-math.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_MODES 8
#define ORC_MAX_SOS 8
#define ORC_N_RAW 21
#define ORC_N_TD 5

typedef struct {
    int32_t fs, n_fft, hop;
    int32_t band_lo, band_hi;                 /* inclusive rfft bins of the operating band */
    int32_t n_modes;
    int32_t mode_lo[ORC_MAX_MODES];           /* inclusive absolute bins; lo > hi == empty */
    int32_t mode_hi[ORC_MAX_MODES];
    int32_t mode_in_band_lo[ORC_MAX_MODES];   /* same masks restricted to the band (classifier) */
    int32_t mode_in_band_hi[ORC_MAX_MODES];
    double mode_weight[ORC_MAX_MODES];
    /* noise-PSD tracker: Python-float constants already cast the way numpy (NEP 50) casts them */
    float trk_eta, trk_scale_alpha, trk_one_minus_alpha, trk_step_floor;
    float trk_q, trk_neg_one_minus_q, trk_maxr;
    double ema_up, ema_down;
    int32_t warmup_need;
    float eps_f32;
    int32_t detector_use_noise_norm;          /* 1: log_sub / ratio_db below; 0: absolute dB */
    int32_t norm_ratio_db;
    /* flux baseline (Python doubles) */
    double bl_q, bl_eta, bl_scale_alpha, bl_floor;
    int32_t norm_enable;
    float norm_min_f32;
    /* decision */
    float thr_primary, thr_m1, thr_m2, thr_m3;
    int32_t min_support;
    float td_gate_thr;
    int32_t has_kurt_upper;
    float kurt_upper;
    float noise_hi, mode_flux_noise_max;
    /* TD prefilter: scipy butter SOS + sosfilt_zi computed by the host wrapper */
    int32_t n_sos, padlen;
    double sos[ORC_MAX_SOS][6];
    double zi[ORC_MAX_SOS][2];
    double eps_f64;
    /* TD block-energy features */
    int32_t blk_len, blk_hop, blk_post_pre, blk_smooth;
    /* raw spectral features */
    int32_t low_lo, low_hi, rain_lo, rain_hi; /* inclusive bins; lo > hi == empty */
    double rolloff_fraction;
    int32_t suppressor_bypass;
    int32_t adaptive_q;                       /* adaptive_q_enable (rain_signal_processor.py:570-576, :634-638) */
    double aq_base, aq_min, aq_alpha;
    int32_t pre_smooth_frames;                /* _time_smooth before the tracker (:366-379, :690-692); <= 1: off */
    int32_t median_frames;                    /* _causal_time_median_filter after it (:381-396, :717-719); <= 1: off */
    int32_t bypass_classifier;                /* every frame NOISE, rain_conf 0 (rain_signal_processor.py:846-857) */
    int32_t reserved;
} orc_params;

typedef struct {
    /* optional planes: NULL == do not export */
    float *S;            /* [T][F][2] complex64 */
    float *P_band;       /* [T][K] */
    float *N1_band;      /* [T][K] detector_noise_psd */
    float *Nlag_band;    /* [T][K] detector_noise_psd_lag */
    float *D_band;       /* [T][K] */
    float *N2_band;      /* [T][K] noise_psd */
    float *mode_flux;    /* [M][T] raw per-mode flux */
    float *flux_modes;   /* [T] */
    float *baseline;     /* [M+1][T]: row 0 total, rows 1..M per mode */
    float *norm_flux;    /* [M][T] */
    float *score;        /* [T] mode_flux_score */
    float *x_td;         /* [N] prefiltered audio */
    float *td;           /* [5][T] crest, kurtosis, blk crest, blk width, blk post/pre */
    float *raw;          /* [21][T] */
    uint8_t *gate;       /* [T] */
    int8_t *frame_class; /* [T] required */
    float *rain_conf;    /* [T] required */
    float *noise_conf;   /* [T] required */
} orc_out;

/* ------------------------------------------------------------------ */
/* numpy / SVML numeric semantics                                       */
/* ------------------------------------------------------------------ */
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* numpy/_core/src/umath/loops_utils.h.src: @TYPE@_pairwise_sum */
static float pw_sum_f32(const float *a, long n, long s)
{
    if (n < 8) {
        float res = -0.0f;
        for (long i = 0; i < n; i++) res += a[i * s];
        return res;
    } else if (n <= 128) {
        float r[8], res;
        long i;
        for (int j = 0; j < 8; j++) r[j] = a[j * s];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[(i + j) * s];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i * s];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return pw_sum_f32(a, n2, s) + pw_sum_f32(a + n2 * s, n - n2, s);
    }
}
static double pw_sum_f64(const double *a, long n, long s)
{
    if (n < 8) {
        double res = -0.0;
        for (long i = 0; i < n; i++) res += a[i * s];
        return res;
    } else if (n <= 128) {
        double r[8], res;
        long i;
        for (int j = 0; j < 8; j++) r[j] = a[j * s];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[(i + j) * s];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i * s];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return pw_sum_f64(a, n2, s) + pw_sum_f64(a + n2 * s, n - n2, s);
    }
}
/* np.add.reduce over a contiguous 1-D run: identity 0 + pairwise over all n elements
 * (verified against np.sum for n = 4..1000 in this container) */
static float np_sum_f32(const float *a, long n)
{
    if (n <= 0) return 0.0f;
    return 0.0f + pw_sum_f32(a, n, 1);
}
static double np_sum_f64(const double *a, long n)
{
    if (n <= 0) return 0.0;
    return 0.0 + pw_sum_f64(a, n, 1);
}

/* numpy complex64 absolute (loops_unary_complex.dispatch.c.src, SIMD path):
 * larger * sqrt(fma(q, q, 1)) with q = smaller / larger.  Verified bit-exact vs np.abs. */
static inline float np_cabsf(float re, float im)
{
    float a = fabsf(re), b = fabsf(im);
    float mx = a > b ? a : b, mn = a > b ? b : a;
    if (mx == 0.0f) return 0.0f;
    if (isinf(mx)) return mx;
    float q = mn / mx;
    return mx * sqrtf(fmaf(q, q, 1.0f));
}

/* numpy/SVML svml_z0_log10_s_la.s (__svml_log10f16), main path, transcribed. */
static const uint32_t SVML_L10_T1[16] = {
    0xbdc9ae9b, 0xbda6fcf4, 0xbd8bac76, 0xbd6bca30, 0xbd48a99b, 0xbd2c0a9f, 0xbd1480db, 0xbd00faf2,
    0xbe823aa9, 0xbe656348, 0xbe4afbb9, 0xbe346895, 0xbe20ffff, 0xbe103a0b, 0xbe01a91c, 0xbde9e84e};
static const uint32_t SVML_L10_T2[16] = {
    0x3e13d888, 0x3e10a87c, 0x3e0b95c3, 0x3e057f0b, 0x3dfde038, 0x3df080d9, 0x3de34c1e, 0x3dd68333,
    0x3dac6e8e, 0x3dd54a51, 0x3df30f40, 0x3e04235d, 0x3e0b7033, 0x3e102c90, 0x3e12ebad, 0x3e141ff8};
static const uint32_t SVML_L10_T3[16] = {
    0xbe5e5a9b, 0xbe5e2677, 0xbe5d83f5, 0xbe5c6016, 0xbe5abd0b, 0xbe58a6fd, 0xbe562e02, 0xbe5362f8,
    0xbe68e27c, 0xbe646747, 0xbe619a73, 0xbe5ff05a, 0xbe5f0570, 0xbe5e92d0, 0xbe5e662b, 0xbe5e5c08};
static const uint32_t SVML_L10_T4[16] = {
    0x3ede5bd8, 0x3ede5b45, 0x3ede57d8, 0x3ede4eb1, 0x3ede3d37, 0x3ede2166, 0x3eddf9d9, 0x3eddc5bb,
    0x3ede08ed, 0x3ede32e7, 0x3ede4967, 0x3ede5490, 0x3ede597f, 0x3ede5b50, 0x3ede5bca, 0x3ede5bd9};

static float svml_log10f(float x)
{
    uint32_t u = f2u(x);
    uint32_t ex = (u >> 23) & 0xff;
    if ((u >> 31) || ex == 0 || ex == 0xff) return log10f(x);  /* rare path: not on our inputs */
    int e = (int)ex - 127;                                     /* vgetexpps(x) */
    uint32_t man = u & 0x7fffff;
    /* vgetmantps imm 0xb: normalise to [0.75, 1.5) */
    uint32_t mb = man | ((man & 0x400000) ? 0x3f000000u : 0x3f800000u);
    float m = u2f(mb);
    int e2 = (man & 0x400000) ? -1 : 0;                        /* vgetexpps(m) */
    uint32_t idx = (mb >> 19) & 0xf;
    float r = m - 1.0f;
    float k = (float)e - (float)e2;
    float p = fmaf(r, u2f(SVML_L10_T1[idx]), u2f(SVML_L10_T2[idx]));
    float kc = k * u2f(0x3e9a209b);
    p = fmaf(r, p, u2f(SVML_L10_T3[idx]));
    p = fmaf(r, p, u2f(SVML_L10_T4[idx]));
    p = fmaf(r, p, kc);
    return p;
}

/* numpy/SVML svml_z0_log1p_s_la.s (__svml_log1pf16), main path, transcribed. */
static float svml_log1pf(float x)
{
    if (!(x > -1.0f) || isinf(x) || isnan(x)) return log1pf(x);
    float A = fmaxf(x, 1.0f), B = fminf(x, 1.0f);
    uint32_t sign = f2u(x) & 0x80000000u;
    float S = A + B;
    uint32_t sb = f2u(S);
    if (sb < 0x00800000u) return log1pf(x);
    int32_t I = (int32_t)(sb - 0x3f2aaaabu);
    float Alo = A - S;
    int32_t N = I >> 23;
    float Rlo = Alo + B;
    float Nf = (float)N;
    float sc = u2f(0x3f800000u - ((uint32_t)N << 23));
    float Rlo_s = Rlo * sc;
    uint32_t M = (uint32_t)I & 0x7fffffu;
    float Mh = u2f(M + 0x3f2aaaabu);
    float R = Mh - 1.0f;
    float r = R + Rlo_s;
    float p = fmaf(u2f(0x3e0d84ed), r, u2f(0xbe1ad9e3));
    p = fmaf(p, r, u2f(0x3e0fcb12));
    p = fmaf(p, r, u2f(0xbe28ad37));
    p = fmaf(p, r, u2f(0x3e4ce190));
    p = fmaf(p, r, u2f(0xbe80058e));
    p = fmaf(p, r, u2f(0x3eaaaa94));
    p = fmaf(p, r, u2f(0xbf000000));
    float q = p * r;
    q = fmaf(q, r, r);
    float res = fmaf(Nf, u2f(0x3f317218), q);
    return u2f(f2u(res) | sign);
}

/* exported for unit pinning against numpy */
void orc_log10f_array(const float *x, float *y, int64_t n) { for (int64_t i = 0; i < n; i++) y[i] = svml_log10f(x[i]); }
void orc_log1pf_array(const float *x, float *y, int64_t n) { for (int64_t i = 0; i < n; i++) y[i] = svml_log1pf(x[i]); }
void orc_cabsf_array(const float *z, float *y, int64_t n) { for (int64_t i = 0; i < n; i++) y[i] = np_cabsf(z[2 * i], z[2 * i + 1]); }
float orc_np_sum_f32(const float *a, int64_t n) { return np_sum_f32(a, n); }

static inline float nan_to_num0(float v) { return (isnan(v) || isinf(v)) ? 0.0f : v; }

/* ------------------------------------------------------------------ */
/* STFT (librosa.stft semantics, float64 FFT rounded to complex64)      */
/* ------------------------------------------------------------------ */
static void fft_c2c_f64(double *re, double *im, int n, const double *twr, const double *twi)
{
    /* iterative radix-2 DIT, bit reversal first */
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            double t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < half; k++) {
                double wr = twr[k * step], wi = twi[k * step];
                double xr = re[i + k + half], xi = im[i + k + half];
                double tr = xr * wr - xi * wi, ti = xr * wi + xi * wr;
                re[i + k + half] = re[i + k] - tr; im[i + k + half] = im[i + k] - ti;
                re[i + k] += tr; im[i + k] += ti;
            }
    }
}

/* scipy.signal.get_window("hann", n, fftbins=True): general_cosine on linspace(-pi, pi, n+1)[:n] */
void orc_hann_periodic(double *w, int n)
{
    double start = -M_PI, step = (M_PI - start) / (double)n;
    for (int i = 0; i < n; i++) {
        double fac = (double)i * step + start;
        w[i] = 0.5 + 0.5 * cos(fac);
    }
}

int64_t orc_num_frames(int64_t n, int hop) { return 1 + n / hop; }
int64_t orc_num_td_frames(int64_t n, int n_fft, int hop) { return n < n_fft ? 0 : 1 + (n - n_fft) / hop; }

/* S: [T][F][2], P: [T][F] */
static void stft_power(const orc_params *p, const double *win, const float *x, int64_t n,
                       int64_t T, float *S, float *P)
{
    const int nfft = p->n_fft, F = nfft / 2 + 1, pad = nfft / 2;
    double *twr = malloc(sizeof(double) * nfft), *twi = malloc(sizeof(double) * nfft);
    double *re = malloc(sizeof(double) * nfft), *im = malloc(sizeof(double) * nfft);
    for (int k = 0; k < nfft; k++) {
        twr[k] = cos(-2.0 * M_PI * (double)k / (double)nfft);
        twi[k] = sin(-2.0 * M_PI * (double)k / (double)nfft);
    }
    for (int64_t t = 0; t < T; t++) {
        int64_t s0 = t * (int64_t)p->hop - pad;
        for (int i = 0; i < nfft; i++) {
            int64_t s = s0 + i;
            double v = (s >= 0 && s < n) ? (double)x[s] : 0.0;
            re[i] = win[i] * v;
            im[i] = 0.0;
        }
        fft_c2c_f64(re, im, nfft, twr, twi);
        for (int k = 0; k < F; k++) {
            float sr = (float)re[k], si = (float)im[k];
            if (k == 0 || k == nfft / 2) si = 0.0f;   /* pocketfft r2c: exactly real DC/Nyquist */
            if (S) { S[(t * F + k) * 2] = sr; S[(t * F + k) * 2 + 1] = si; }
            float a = np_cabsf(sr, si);
            P[t * F + k] = a * a;
        }
    }
    free(twr); free(twi); free(re); free(im);
}

/* ------------------------------------------------------------------ */
/* noise-PSD tracker, one pass (rain_signal_processor.py:555-721)       */
/* exclude[t] != 0  <=> is_rain_for_psd[t]                              */
/* P: [T][F]; N out: [T][K]                                             */
/* ------------------------------------------------------------------ */
static int cmp_f32(const void *a, const void *b) { float x = *(const float *)a, y = *(const float *)b; return (x > y) - (x < y); }

static void track_noise_psd(const orc_params *p, const float *P_in, int64_t T, int F_in,
                            const uint8_t *exclude, float *N)
{
    const int K = p->band_hi - p->band_lo + 1;
    const float *P = P_in;
    int F = F_in;
    float *Y = NULL;
    if (p->pre_smooth_frames > 1) {
        /* moving average from a float32 cumulative sum over time (np.cumsum is sequential), band bins only */
        const int L = p->pre_smooth_frames;
        float *cs = malloc(sizeof(float) * T * K);
        Y = malloc(sizeof(float) * T * K);
        for (int k = 0; k < K; k++) {
            float acc = 0.0f;
            for (int64_t t = 0; t < T; t++) {
                acc = (t == 0) ? P_in[p->band_lo + k] : acc + P_in[t * F_in + p->band_lo + k];
                cs[t * K + k] = acc;
                int64_t t0 = t - L + 1 > 0 ? t - L + 1 : 0;
                Y[t * K + k] = (t0 == 0) ? cs[t * K + k] / (float)(t + 1)
                                         : (cs[t * K + k] - cs[(t0 - 1) * K + k]) / (float)(t - t0 + 1);
            }
        }
        free(cs);
        P = Y - p->band_lo;      /* rows of K values: the code below reads P[t * F + band_lo + k] */
        F = K;
    }
    float *trk = malloc(sizeof(float) * K), *ts = malloc(sizeof(float) * K);
    int warm = 0;
    double rain_ema = 0.0;                    /* rain_prev_ema: EMA of the excluded-frame flags, a Python float */
    for (int k = 0; k < K; k++) {
        float p0 = P[p->band_lo + k];
        trk[k] = fmaxf(p0, 0.0f);
        ts[k] = fmaxf(fabsf(p0), p->trk_step_floor);
    }
    for (int64_t t = 0; t < T; t++) {
        const float *Pt = P + t * F + p->band_lo;
        float *Nt = N + t * K;
        int excl = exclude ? exclude[t] != 0 : 0;
        int allow = (warm < p->warmup_need) || !excl;
        if (t == 0) {
            if (allow) warm++;
            for (int k = 0; k < K; k++) {
                float nb = fminf(trk[k], p->trk_maxr * Pt[k]);
                Nt[k] = fmaxf(nb, 0.0f);
            }
            rain_ema = p->aq_alpha * rain_ema + (1.0 - p->aq_alpha) * (excl ? 1.0 : 0.0);
            continue;
        }
        const float *Np = N + (t - 1) * K;
        float qf = p->trk_q, nqf = p->trk_neg_one_minus_q;
        if (p->adaptive_q) {                  /* q_eff is a Python float; numpy casts it to float32 at the product */
            double q = p->aq_base - (p->aq_base - p->aq_min) * rain_ema;
            q = q < p->aq_min ? p->aq_min : (q > p->aq_base ? p->aq_base : q);
            qf = (float)q; nqf = (float)(-(1.0 - q));
        }
        for (int k = 0; k < K; k++) {
            float pk = Pt[k];
            float err = pk - trk[k];
            ts[k] = p->trk_scale_alpha * ts[k] + p->trk_one_minus_alpha * fabsf(err);
            float step = p->trk_eta * fmaxf(ts[k], p->trk_step_floor);
            float delta = (pk >= trk[k]) ? qf * step : nqf * step;
            float cand = fmaxf(trk[k] + delta, 0.0f);
            if (allow) trk[k] = cand;
            float raw = trk[k];
            double lam = (raw > Np[k]) ? p->ema_up : p->ema_down;
            double nb = lam * (double)Np[k] + (1.0 - lam) * (double)raw;
            double cap = (double)(p->trk_maxr * pk);
            if (cap < nb) nb = cap;          /* np.minimum */
            if (nb < 0.0) nb = 0.0;          /* np.maximum(.,0) */
            Nt[k] = (float)nb;
        }
        if (allow) warm++;
        rain_ema = p->aq_alpha * rain_ema + (1.0 - p->aq_alpha) * (excl ? 1.0 : 0.0);
    }
    free(trk); free(ts);
    free(Y);
    if (p->median_frames > 1) {
        /* causal median over time per bin, window [max(0, t - L + 1), t], L made odd; np.median: even counts average
         * the two middle values in float32 */
        int L = p->median_frames;
        if (L % 2 == 0) L += 1;
        float *R = malloc(sizeof(float) * T * K), *w = malloc(sizeof(float) * L);
        memcpy(R, N, sizeof(float) * T * K);
        for (int64_t t = 0; t < T; t++) {
            int64_t t0 = t - L + 1 > 0 ? t - L + 1 : 0;
            int n = (int)(t - t0 + 1);
            for (int k = 0; k < K; k++) {
                for (int i = 0; i < n; i++) w[i] = R[(t0 + i) * K + k];
                qsort(w, n, sizeof(float), cmp_f32);
                N[t * K + k] = (n & 1) ? w[n / 2] : (w[n / 2 - 1] + w[n / 2]) / 2.0f;
            }
        }
        free(R); free(w);
    }
}

/* ------------------------------------------------------------------ */
/* scipy.signal.sosfiltfilt (padtype="odd")                             */
/* ------------------------------------------------------------------ */
static void sosfilt_inplace(const orc_params *p, double *x, int64_t n, double z[][2])
{
    const int ns = p->n_sos;
    for (int64_t i = 0; i < n; i++) {
        double xc = x[i];
        for (int s = 0; s < ns; s++) {
            const double *c = p->sos[s];
            double xn = c[0] * xc + z[s][0];
            z[s][0] = (c[1] * xc - c[4] * xn) + z[s][1];
            z[s][1] = c[2] * xc - c[5] * xn;
            xc = xn;
        }
        x[i] = xc;
    }
}

static int sosfiltfilt_f32(const orc_params *p, const float *x, int64_t n, float *y)
{
    const int edge = p->padlen;
    if (p->n_sos == 0) { memcpy(y, x, sizeof(float) * n); return 0; }
    if (n <= edge) return -1;
    int64_t m = n + 2 * edge;
    double *ext = malloc(sizeof(double) * m);
    double z[ORC_MAX_SOS][2];
    for (int i = 0; i < edge; i++) ext[i] = 2.0 * (double)x[0] - (double)x[edge - i];
    for (int64_t i = 0; i < n; i++) ext[edge + i] = (double)x[i];
    for (int i = 0; i < edge; i++) ext[edge + n + i] = 2.0 * (double)x[n - 1] - (double)x[n - 2 - i];
    for (int s = 0; s < p->n_sos; s++) { z[s][0] = p->zi[s][0] * ext[0]; z[s][1] = p->zi[s][1] * ext[0]; }
    sosfilt_inplace(p, ext, m, z);
    for (int64_t i = 0; i < m / 2; i++) { double t = ext[i]; ext[i] = ext[m - 1 - i]; ext[m - 1 - i] = t; }
    for (int s = 0; s < p->n_sos; s++) { z[s][0] = p->zi[s][0] * ext[0]; z[s][1] = p->zi[s][1] * ext[0]; }
    sosfilt_inplace(p, ext, m, z);
    for (int64_t i = 0; i < n; i++) y[i] = (float)ext[m - 1 - edge - i];
    free(ext);
    return 0;
}

/* ------------------------------------------------------------------ */
/* TD features (feature_extraction.py:174-538, td_input_mode="default") */
/* td: [5][T] zero-filled; Tloc frames are written                      */
/* ------------------------------------------------------------------ */
static double peak_width_half(const double *x, int n, int peak)
{
    /* scipy.signal.peak_prominences (wlen=None) + peak_widths(rel_height=0.5) */
    int i = peak, lb = peak, rb = peak;
    double lmin = x[peak], rmin = x[peak];
    while (0 <= i && x[i] <= x[peak]) { if (x[i] < lmin) { lmin = x[i]; lb = i; } i--; }
    i = peak;
    while (i <= n - 1 && x[i] <= x[peak]) { if (x[i] < rmin) { rmin = x[i]; rb = i; } i++; }
    double prom = x[peak] - (lmin > rmin ? lmin : rmin);
    double height = x[peak] - prom * 0.5;
    i = peak;
    while (lb < i && height < x[i]) i--;
    double lip = (double)i;
    if (x[i] < height) lip += (height - x[i]) / (x[i + 1] - x[i]);
    i = peak;
    while (i < rb && height < x[i]) i++;
    double rip = (double)i;
    if (x[i] < height) rip -= (height - x[i]) / (x[i - 1] - x[i]);
    return rip - lip;
}

static void td_features(const orc_params *p, const float *xt, int64_t n, int64_t T, float *td)
{
    const int L = p->n_fft, hop = p->hop;
    const double eps = p->eps_f64;
    int64_t Tloc = orc_num_td_frames(n, L, hop);
    float *crest = td, *kurt = td + T, *bcrest = td + 2 * T, *bwidth = td + 3 * T, *bratio = td + 4 * T;
    memset(td, 0, sizeof(float) * 5 * T);
    if (Tloc > T) Tloc = T;
    float *sq = malloc(sizeof(float) * L), *d2 = malloc(sizeof(float) * L), *d4 = malloc(sizeof(float) * L);
    for (int64_t t = 0; t < Tloc; t++) {
        const float *seg = xt + t * hop;
        float pk = 0.0f;
        for (int i = 0; i < L; i++) { sq[i] = seg[i] * seg[i]; float a = fabsf(seg[i]); if (a > pk) pk = a; }
        float mean_sq = np_sum_f32(sq, L) / (float)L;
        float rms = sqrtf(mean_sq + (float)eps);
        double r = (double)rms;
        crest[t] = (float)((double)pk / (r > eps ? r : eps));
        if (L >= 4) {
            float mean = np_sum_f32(seg, L) / (float)L;
            /* scipy _moment: (a-mean)**order with a float32 0-d exponent -> numpy SVML powf for order 4
             * (89% correctly rounded); restated as the correctly rounded 4th power.  Tolerance feature. */
            for (int i = 0; i < L; i++) {
                float d = seg[i] - mean; d2[i] = d * d;
                double dd = (double)d; d4[i] = (float)((dd * dd) * (dd * dd));
            }
            float m2 = np_sum_f32(d2, L) / (float)L, m4 = np_sum_f32(d4, L) / (float)L;
            float lim = 1.1920929e-07f * mean;
            float kv;
            if (m2 <= lim * lim) kv = NAN;
            else {
                double nn = (double)L;
                float a = ((float)(nn * nn - 1.0) * m4) / (m2 * m2) - (float)(3.0 * (nn - 1.0) * (nn - 1.0));
                kv = (float)(1.0 / (nn - 2.0) / (nn - 3.0)) * a + 3.0f;
            }
            kurt[t] = (isnan(kv) || isinf(kv)) ? 0.0f : kv;
        }
    }
    free(sq); free(d2); free(d4);

    /* block-energy envelope over the whole clip */
    const int B = p->blk_len > 1 ? p->blk_len : 1;
    const int H = p->blk_hop > 1 ? p->blk_hop : 1;
    if (n >= B) {
        int64_t nb = (n - B) / H + 1;
        double *csum = malloc(sizeof(double) * (n + 1));
        double *env = malloc(sizeof(double) * nb), *env2 = malloc(sizeof(double) * nb);
        csum[0] = 0.0;
        for (int64_t i = 0; i < n; i++) { double v = (double)xt[i]; csum[i + 1] = csum[i] + v * v; }
        for (int64_t b = 0; b < nb; b++) {
            double s = csum[b * H + B] - csum[b * H];
            double e = s / (double)B;
            env[b] = sqrt(e > 0.0 ? e : 0.0);
        }
        if (p->blk_smooth && nb >= 3) {
            for (int64_t b = 0; b < nb; b++) {
                double acc = 0.0;
                if (b > 0) acc = env[b - 1] * 0.25;
                acc = (b > 0) ? acc + env[b] * 0.5 : env[b] * 0.5;
                if (b + 1 < nb) acc += env[b + 1] * 0.25;
                env2[b] = acc;
            }
            double *tmp = env; env = env2; env2 = tmp;
        }
        int bpf = (L + H - 1) / H; if (bpf < 1) bpf = 1;
        int bstep = (int)nearbyint((double)hop / (double)H); if (bstep < 1) bstep = 1;
        int pp = p->blk_post_pre > 1 ? p->blk_post_pre : 1;
        double *fe2 = malloc(sizeof(double) * bpf);
        for (int64_t t = 0; t < Tloc; t++) {
            int64_t b0 = t * bstep, b1 = b0 + bpf;
            if (b1 > nb) b1 = nb;
            if (b1 <= b0) continue;
            int m = (int)(b1 - b0);
            const double *fe = env + b0;
            int pi = 0;
            for (int i = 0; i < m; i++) { fe2[i] = fe[i] * fe[i]; if (fe[i] > fe[pi]) pi = i; }
            double rms = sqrt(np_sum_f64(fe2, m) / (double)m);
            double pv = fe[pi];
            bcrest[t] = (float)(pv / (rms > eps ? rms : eps));
            float w = 0.0f;
            if (pv > eps && m >= 3 && pi > 0 && pi < m - 1) {
                double lv = fe[pi - 1], rv = fe[pi + 1];
                double prom = pv - (lv > rv ? lv : rv);
                if (prom > eps) {
                    double wv = peak_width_half(fe, m, pi);
                    if (isfinite(wv) && wv > 0.0) w = (float)wv;
                }
            }
            bwidth[t] = w;
            int64_t pidx = b0 + pi;
            int64_t pre0 = pidx - pp > 0 ? pidx - pp : 0, pre1 = pidx;
            int64_t po0 = pidx + 1, po1 = pidx + 1 + pp < nb ? pidx + 1 + pp : nb;
            double pre = pre1 > pre0 ? np_sum_f64(env + pre0, pre1 - pre0) / (double)(pre1 - pre0) : 0.0;
            double post = po1 > po0 ? np_sum_f64(env + po0, po1 - po0) / (double)(po1 - po0) : 0.0;
            bratio[t] = (float)log((post + eps) / (pre + eps));
        }
        free(fe2); free(csum); free(env); free(env2);
    }
    for (int64_t i = 0; i < 5 * T; i++) td[i] = nan_to_num0(td[i]);
}

/* ------------------------------------------------------------------ */
/* raw spectral features (feature_extraction.py:542-747), float64       */
/* ------------------------------------------------------------------ */
static double sum_bins(const float *Pt, int lo, int hi)
{
    double s = 0.0;
    for (int k = lo; k <= hi; k++) s += (double)Pt[k];
    return s;
}

static void raw_features(const orc_params *p, const float *freqs, const float *P, int64_t T, int F, float *raw)
{
    const double eps = p->eps_f64;
    const int lo = p->band_lo, hi = p->band_hi, K = hi - lo + 1;
    const int ncep = 2 * (K - 1);
    double *lg = malloc(sizeof(double) * K);
    for (int64_t t = 0; t < T; t++) {
        const float *Pt = P + t * F;
        double total = sum_bins(Pt, 0, F - 1) + eps;
        double total_nodc = F > 1 ? sum_bins(Pt, 1, F - 1) + eps : total;
        double op = sum_bins(Pt, lo, hi) + eps;
        double shape_total = op;
        double cen = 0.0;
        for (int k = lo; k <= hi; k++) cen += (double)freqs[k] * (double)Pt[k];
        cen /= shape_total;
        double bw = 0.0;
        for (int k = lo; k <= hi; k++) { double d = (double)freqs[k] - cen; bw += d * d * (double)Pt[k]; }
        bw = sqrt(bw / shape_total);
        double lowr = p->low_lo <= p->low_hi ? sum_bins(Pt, p->low_lo, p->low_hi) / total_nodc : 0.0;
        double rainr = p->rain_lo <= p->rain_hi ? sum_bins(Pt, p->rain_lo, p->rain_hi) / total_nodc : 0.0;
        double mbp[ORC_MAX_MODES], mtot = 0.0, ratio[ORC_MAX_MODES];
        for (int i = 0; i < p->n_modes; i++) {
            mbp[i] = p->mode_lo[i] <= p->mode_hi[i] ? sum_bins(Pt, p->mode_lo[i], p->mode_hi[i]) : 0.0;
            mtot += mbp[i];
        }
        mtot += eps;
        double ent = 0.0, mean = 0.0, mx = -INFINITY;
        for (int i = 0; i < p->n_modes; i++) {
            ratio[i] = mbp[i] / mtot;
            ent += ratio[i] * log(ratio[i] + eps);
            mean += ratio[i];
            if (ratio[i] > mx) mx = ratio[i];
        }
        ent = -ent;
        mean /= (double)p->n_modes;
        double var = 0.0;
        for (int i = 0; i < p->n_modes; i++) { double d = ratio[i] - mean; var += d * d; }
        double sd = sqrt(var / (double)p->n_modes);
        double mlog = 0.0, marith = 0.0;
        for (int k = lo; k <= hi; k++) { mlog += log((double)Pt[k] + eps); marith += (double)Pt[k] + eps; }
        double flat = exp(mlog / (double)K) / (marith / (double)K + eps);
        double frac = p->rolloff_fraction < 0.0 ? 0.0 : (p->rolloff_fraction > 1.0 ? 1.0 : p->rolloff_fraction);
        double thr = frac * shape_total, cs = 0.0;
        int ridx = 0, found = 0, dom = 0;
        for (int k = 0; k < K; k++) {
            cs += (double)Pt[lo + k];
            if (!found && cs >= thr) { ridx = k; found = 1; }
            if (Pt[lo + k] > Pt[lo + dom]) dom = k;
        }
        for (int k = 0; k < K; k++) { double v = (double)Pt[lo + k]; lg[k] = log(v > eps ? v : eps); }
        double cep[5] = {0, 0, 0, 0, 0};
        if (K >= 2)
            for (int j = 0; j < 5 && j < ncep; j++) {
                double acc = lg[0] + ((j & 1) ? -lg[K - 1] : lg[K - 1]);
                for (int k = 1; k < K - 1; k++) acc += 2.0 * lg[k] * cos(2.0 * M_PI * (double)j * (double)k / (double)ncep);
                cep[j] = acc / (double)ncep;
            }
        float *o = raw + t;
        o[0 * T] = (float)cen; o[1 * T] = (float)bw; o[2 * T] = (float)lowr; o[3 * T] = (float)rainr;
        for (int i = 0; i < 5; i++) o[(4 + i) * T] = i < p->n_modes ? (float)ratio[i] : 0.0f;
        o[9 * T] = (float)ent; o[10 * T] = (float)sd; o[11 * T] = (float)mx; o[12 * T] = (float)flat;
        o[13 * T] = freqs[lo + ridx]; o[14 * T] = freqs[lo + dom]; o[15 * T] = (float)op;
        for (int j = 0; j < 5; j++) o[(16 + j) * T] = (float)cep[j];
    }
    free(lg);
}

/* ------------------------------------------------------------------ */
/* causal_stochastic_low_quantile_baseline (rain_frame_classifier.py:31-82) */
/* ------------------------------------------------------------------ */
void orc_baseline(const float *x, int64_t T, double q, double eta, double scale_alpha, double floor_, float *out)
{
    if (T <= 0) return;
    double x0 = (double)x[0];
    double baseline = x0 > floor_ ? x0 : floor_;
    double scale = fabs(x0) > floor_ ? fabs(x0) : floor_;
    for (int64_t t = 0; t < T; t++) {
        out[t] = (float)baseline;
        double xt = (double)x[t];
        double err = xt - baseline;
        scale = scale_alpha * scale + (1.0 - scale_alpha) * fabs(err);
        double step = eta * (scale > floor_ ? scale : floor_);
        double delta = (xt >= baseline) ? q * step : -(1.0 - q) * step;
        double nb = baseline + delta;
        baseline = nb > floor_ ? nb : floor_;
    }
    float ff = (float)floor_;
    for (int64_t t = 0; t < T; t++) {
        float v = out[t];
        if (isnan(v) || isinf(v)) v = ff;
        out[t] = v > ff ? v : ff;
    }
}

/* ------------------------------------------------------------------ */
/* full pipeline for one clip (SpectralNoiseProcessor.process, default flags) */
/* ------------------------------------------------------------------ */
int orc_process(const orc_params *p, const double *window, const float *freqs,
                const float *x, int64_t n, orc_out *o)
{
    const int F = p->n_fft / 2 + 1, K = p->band_hi - p->band_lo + 1, M = p->n_modes;
    const int64_t T = orc_num_frames(n, p->hop);
    if (K <= 0 || M < 4 || M > ORC_MAX_MODES) return -2;
    float *P = malloc(sizeof(float) * T * F);
    stft_power(p, window, x, n, T, o->S, P);
    if (o->P_band)
        for (int64_t t = 0; t < T; t++) memcpy(o->P_band + t * K, P + t * F + p->band_lo, sizeof(float) * K);

    /* detector normalisation */
    float *N1 = malloc(sizeof(float) * T * K), *D = malloc(sizeof(float) * T * K);
    if (p->detector_use_noise_norm) {
        track_noise_psd(p, P, T, F, NULL, N1);
        for (int64_t t = 0; t < T; t++)
            for (int k = 0; k < K; k++) {
                float pk = P[t * F + p->band_lo + k];
                float nl = N1[(t > 0 ? t - 1 : 0) * K + k];
                nl = fminf(nl, p->trk_maxr * pk);
                if (o->Nlag_band) o->Nlag_band[t * K + k] = nl;
                if (p->norm_ratio_db)
                    D[t * K + k] = 10.0f * svml_log10f(pk / (nl + p->eps_f32) + p->eps_f32);
                else
                    D[t * K + k] = 10.0f * svml_log10f(pk + p->eps_f32) - 10.0f * svml_log10f(nl + p->eps_f32);
            }
        if (o->N1_band) memcpy(o->N1_band, N1, sizeof(float) * T * K);
    } else {
        for (int64_t t = 0; t < T; t++)
            for (int k = 0; k < K; k++) D[t * K + k] = 10.0f * svml_log10f(P[t * F + p->band_lo + k] + p->eps_f32);
    }
    if (o->D_band) memcpy(o->D_band, D, sizeof(float) * T * K);

    /* TD features on the zero-phase prefiltered waveform */
    float *xt = malloc(sizeof(float) * (n > 0 ? n : 1));
    float *td = malloc(sizeof(float) * 5 * T);
    if (sosfiltfilt_f32(p, x, n, xt) != 0) { free(P); free(N1); free(D); free(xt); free(td); return -3; }
    td_features(p, xt, n, T, td);
    if (o->x_td) memcpy(o->x_td, xt, sizeof(float) * n);
    if (o->td) memcpy(o->td, td, sizeof(float) * 5 * T);
    if (o->raw) raw_features(p, freqs, P, T, F, o->raw);

    /* flux (rain_frame_classifier.py:713-759) */
    float *mf = malloc(sizeof(float) * M * T), *fm = malloc(sizeof(float) * T);
    float *flux = malloc(sizeof(float) * K);
    for (int64_t t = 0; t < T; t++) {
        if (t < 2) {
            for (int i = 0; i < M; i++) mf[i * T + t] = 0.0f;
            fm[t] = 0.0f;
            continue;
        }
        for (int k = 0; k < K; k++) {
            float d = D[t * K + k] - D[(t - 2) * K + k];
            flux[k] = d > 0.0f ? d : 0.0f;     /* np.maximum(delta2, 0.0) */
            if (isnan(d)) flux[k] = d;
        }
        double tot = 0.0;
        for (int i = 0; i < M; i++) {
            int a = p->mode_in_band_lo[i], b = p->mode_in_band_hi[i];
            float s = a <= b ? np_sum_f32(flux + a, b - a + 1) : 0.0f;
            mf[i * T + t] = s;
            tot += p->mode_weight[i] * (double)s;
        }
        fm[t] = (float)tot;
    }
    free(flux);
    if (o->mode_flux) memcpy(o->mode_flux, mf, sizeof(float) * M * T);
    if (o->flux_modes) memcpy(o->flux_modes, fm, sizeof(float) * T);

    /* baselines + normalisation (rain_frame_classifier.py:873-893) */
    float *bl = malloc(sizeof(float) * T), *score = malloc(sizeof(float) * T), *nf = malloc(sizeof(float) * M * T);
    orc_baseline(fm, T, p->bl_q, p->bl_eta, p->bl_scale_alpha, p->bl_floor, bl);
    if (o->baseline) memcpy(o->baseline, bl, sizeof(float) * T);
    for (int64_t t = 0; t < T; t++) {
        float ex = fmaxf(fm[t] - bl[t], 0.0f);
        float sc = p->norm_enable ? ex / (bl[t] + p->norm_min_f32) : ex;
        score[t] = nan_to_num0(sc);
    }
    for (int i = 0; i < M; i++) {
        orc_baseline(mf + i * T, T, p->bl_q, p->bl_eta, p->bl_scale_alpha, p->bl_floor, bl);
        if (o->baseline) memcpy(o->baseline + (i + 1) * T, bl, sizeof(float) * T);
        for (int64_t t = 0; t < T; t++) {
            float ex = fmaxf(mf[i * T + t] - bl[t], 0.0f);
            float sc = p->norm_enable ? ex / (bl[t] + p->norm_min_f32) : ex;
            nf[i * T + t] = nan_to_num0(sc);
        }
    }
    if (o->score) memcpy(o->score, score, sizeof(float) * T);
    if (o->norm_flux) memcpy(o->norm_flux, nf, sizeof(float) * M * T);

    /* gate + decision + labels (rain_frame_classifier.py:230-284, 937-998) */
    uint8_t *excl = malloc(T > 0 ? T : 1);
    for (int64_t t = 0; t < T; t++) {
        float crest = td[t], kurt = td[T + t];
        int gate = crest > p->td_gate_thr;
        if (p->has_kurt_upper) gate = gate && (kurt <= p->kurt_upper);
        float g = gate ? 1.0f : 0.0f;
        float f0 = svml_log1pf(fmaxf(nf[0 * T + t] * g, 0.0f));
        float f1 = svml_log1pf(fmaxf(nf[1 * T + t] * g, 0.0f));
        float f2 = svml_log1pf(fmaxf(nf[2 * T + t] * g, 0.0f));
        float f3 = svml_log1pf(fmaxf(nf[3 * T + t] * g, 0.0f));
        int hits = (f1 >= p->thr_m1) + (f2 >= p->thr_m2) + (f3 >= p->thr_m3);
        int ms = p->min_support > 1 ? p->min_support : 1;
        int is_rain = (f0 >= p->thr_primary) && (hits >= ms);
        float rc = is_rain ? 1.0f : 0.0f;
        float nc = 1.0f - rc; nc = nc < 0.0f ? 0.0f : (nc > 1.0f ? 1.0f : nc);
        int weak = (score[t] * g) <= p->mode_flux_noise_max;
        int8_t cls = 1;
        if (nc >= p->noise_hi && weak && !is_rain) cls = 0;
        if (is_rain) cls = 2;
        if (p->bypass_classifier) { cls = 0; rc = 0.0f; nc = 1.0f; }
        o->frame_class[t] = cls; o->rain_conf[t] = rc; o->noise_conf[t] = nc;
        if (o->gate) o->gate[t] = (uint8_t)gate;
        excl[t] = cls != 0;                     /* is_rain_for_psd = ~is_noise */
    }

    /* final noise PSD (rain_signal_processor.py:1028) */
    if (o->N2_band) {
        if (p->suppressor_bypass) memset(o->N2_band, 0, sizeof(float) * T * K);
        else track_noise_psd(p, P, T, F, excl, o->N2_band);
    }
    free(P); free(N1); free(D); free(xt); free(td); free(mf); free(fm); free(bl); free(score); free(nf); free(excl);
    return 0;
}

/* dB noise floor of a [T][K] PSD: 10*log10(N + eps) with numpy float32 semantics */
void orc_noise_db(const float *N, int64_t n, float eps, float *db)
{
    for (int64_t i = 0; i < n; i++) db[i] = 10.0f * svml_log10f(N[i] + eps);
}

/* safe_to_float for int16 PCM (audio_io.py:71-72) */
void orc_pcm_to_f32(const int16_t *pcm, int64_t n, float *x)
{
    for (int64_t i = 0; i < n; i++) x[i] = (float)pcm[i] / 32767.0f;
}

/* Batch driver used only for the CPU baseline timing: clips are independent units; a pthread
 * pool pulls clip indices from a shared counter (the analogue of the reference's
 * ProcessPoolExecutor over files, audio_processing_framework.py:249-290).  N2 is tracked (and
 * discarded) so the work matches the full pipeline. */
#include <pthread.h>
typedef struct {
    const orc_params *p; const double *window; const float *freqs;
    const int16_t *pcm; const int64_t *offsets; int n_clips;
    int8_t *frame_class; const int64_t *frame_offsets; int32_t *rain_count;
    int next; int rc; pthread_mutex_t mu;
} orc_batch;

static void *orc_batch_worker(void *arg)
{
    orc_batch *b = (orc_batch *)arg;
    for (;;) {
        pthread_mutex_lock(&b->mu);
        int c = b->next++;
        pthread_mutex_unlock(&b->mu);
        if (c >= b->n_clips) break;
        const orc_params *p = b->p;
        int64_t n = b->offsets[c + 1] - b->offsets[c];
        int64_t T = orc_num_frames(n, p->hop);
        int K = p->band_hi - p->band_lo + 1;
        float *x = malloc(sizeof(float) * (n > 0 ? n : 1));
        orc_pcm_to_f32(b->pcm + b->offsets[c], n, x);
        orc_out o; memset(&o, 0, sizeof(o));
        float *rc = malloc(sizeof(float) * T), *nc = malloc(sizeof(float) * T);
        float *N2 = malloc(sizeof(float) * T * K);
        o.frame_class = b->frame_class + b->frame_offsets[c]; o.rain_conf = rc; o.noise_conf = nc; o.N2_band = N2;
        int r = orc_process(p, b->window, b->freqs, x, n, &o);
        int32_t cnt = 0;
        for (int64_t t = 0; t < T; t++) cnt += o.frame_class[t] == 2;
        b->rain_count[c] = cnt;
        if (r != 0) { pthread_mutex_lock(&b->mu); b->rc = r; pthread_mutex_unlock(&b->mu); }
        free(x); free(rc); free(nc); free(N2);
    }
    return NULL;
}

int orc_process_batch_i16(const orc_params *p, const double *window, const float *freqs,
                          const int16_t *pcm, const int64_t *offsets, int n_clips,
                          int8_t *frame_class, const int64_t *frame_offsets, int32_t *rain_count,
                          int n_threads)
{
    orc_batch b;
    memset(&b, 0, sizeof(b));
    b.p = p; b.window = window; b.freqs = freqs; b.pcm = pcm; b.offsets = offsets; b.n_clips = n_clips;
    b.frame_class = frame_class; b.frame_offsets = frame_offsets; b.rain_count = rain_count;
    pthread_mutex_init(&b.mu, NULL);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_clips) n_threads = n_clips > 0 ? n_clips : 1;
    pthread_t *th = malloc(sizeof(pthread_t) * n_threads);
    for (int i = 1; i < n_threads; i++) pthread_create(&th[i], NULL, orc_batch_worker, &b);
    orc_batch_worker(&b);
    for (int i = 1; i < n_threads; i++) pthread_join(th[i], NULL);
    free(th);
    pthread_mutex_destroy(&b.mu);
    return b.rc;
}

int orc_sizeof_params(void) { return (int)sizeof(orc_params); }
