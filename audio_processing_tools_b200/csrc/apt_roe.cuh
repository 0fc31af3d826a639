// Legacy "RoE" rain detector on the GPU (reference edge/dsp_rain_detection.py:2230-2731).
//
// The reference walks a clip in 2-second parts (:2603-2636); each part is filtered, transformed and scored on its
// own, so the unit of parallelism is the part:
//   roe_filter_kernel   one warp per part: the order-8 input band-pass and, chained behind it, the order-4
//                       400-900 Hz band-pass of the time-domain branch run as a 12-lane systolic cascade (lane =
//                       biquad section, samples move one lane per step through shuffles) -- the arithmetic per
//                       section is scipy's sosfilt recursion, from a zero state per part like the reference
//   roe_frame_kernel    32 frames per CTA: float64 rFFT-256 magnitudes (centred, zero padded, librosa.stft),
//                       kurtosis / crest factor of the frame, energy of the 400-900 Hz signal (:657-767)
//   roe_part_kernel     one CTA per part, thread per frame: band-limited spectral novelty, local-average SNR,
//                       find_peaks mask, thresholding (:1892-1955), peak-in-band gating (:1649-1698), the
//                       estimated natural frequency and the five harmonic novelties
//   roe_state_kernel    one thread: the `max_harmonics` value each part sees (module state of the reference that
//                       survives from part to part, clip to clip and call to call, :1141 :1394-1403)
//   roe_combine_kernel  one CTA per part: harmonic sum, clipping, rain status per frame, energy rise / minimum,
//                       time-domain peaks (:2503-2531, :770-801)
//   roe_clip_kernel     thread per clip: drop counting over the parts, FP / FN combination (:2638-2731)
#pragma once
#include <cuda_runtime.h>
#include "apt_math.cuh"
#include "../../include/apt_b200.h"

namespace apt {

typedef apt_roe_params_t RoeDev;
constexpr int ROE_NT = 256;
constexpr int ROE_LMAX = 256;      // frames + 1 of a part (2 s at hop 128: 176)
constexpr int ROE_FO = APT_ROE_FRAME_F;
constexpr int ROE_PO = APT_ROE_PART_F;

struct RoeParts {
    int n_parts;
    const int32_t* clip;       // [n_parts] clip of the part (ascending)
    const int64_t* start;      // [n_parts] first sample, absolute index into the PCM buffer
    const int32_t* len;        // [n_parts] samples
    const int64_t* fo;         // [n_parts + 1] prefix of (frames + 1)
    const int64_t* yo;         // [n_parts + 1] prefix of len (filtered signal), the 400-900 Hz signal sits at yo + 256 * part
};

__device__ __forceinline__ double roe_load(const int16_t* p, int64_t i) { return (double)pcm_to_f32(__ldg(p + i)); }
__device__ __forceinline__ double roe_load(const float* p, int64_t i) { return (double)__ldg(p + i); }
__device__ __forceinline__ double shfl64(double v, int src) {
    return __hiloint2double(__shfl_sync(0xffffffffu, __double2hiint(v), src), __shfl_sync(0xffffffffu, __double2loint(v), src));
}

// Python's float floor division a // b (b > 0)
__device__ __forceinline__ double py_floordiv(double a, double b) {
    double mod = fmod(a, b);
    double div = (a - mod) / b;
    if (mod != 0.0 && mod < 0.0) { mod += b; div -= 1.0; }
    if (div != 0.0) {
        double fl = floor(div);
        if (div - fl > 0.5) fl += 1.0;
        return fl;
    }
    return 0.0;
}

template <typename PCM>
__global__ void __launch_bounds__(32) roe_filter_kernel(const __grid_constant__ RoeDev p, RoeParts pt, const PCM* __restrict__ pcm,
                                                        double* __restrict__ ybuf, double* __restrict__ tbuf) {
    const int part = blockIdx.x, lane = threadIdx.x;
    const int len = pt.len[part];
    const int64_t g0 = pt.start[part];
    double* y = ybuf + pt.yo[part];
    double* f2 = tbuf + pt.yo[part] + (int64_t)256 * part;      // len + 256 values: zeros(128), y, zeros(128) filtered
    const int ns1 = p.ns_in, ns2 = p.want_td ? p.ns_td : 0, nl = ns1 + ns2;
    double b0 = 0, b1 = 0, b2 = 0, a1 = 0, a2 = 0;
    if (lane < ns1) { b0 = p.sos_in[lane][0]; b1 = p.sos_in[lane][1]; b2 = p.sos_in[lane][2]; a1 = p.sos_in[lane][4]; a2 = p.sos_in[lane][5]; }
    else if (lane < nl) { const int s = lane - ns1; b0 = p.sos_td[s][0]; b1 = p.sos_td[s][1]; b2 = p.sos_td[s][2]; a1 = p.sos_td[s][4]; a2 = p.sos_td[s][5]; }
    for (int i = lane; i < 128 && ns2 > 0; i += 32) f2[i] = 0.0;     // a filter at rest stays at rest on the leading zeros
    double z0 = 0.0, z1 = 0.0, out = 0.0;
    const int steps = len + 128 + nl;
    for (int t0 = 0; t0 < steps; t0 += 32) {
        const double xin = (t0 + lane < len) ? roe_load(pcm, g0 + t0 + lane) : 0.0;
        double ykeep = 0.0, fkeep = 0.0;
        for (int i = 0; i < 32; i++) {
            const int t = t0 + i;
            const double prev = shfl64(out, lane - 1);     // output of the previous section at step t - 1
            const double xi = shfl64(xin, i);              // every lane takes part in every shuffle
            double x = (lane == 0) ? xi : prev;
            // the time-domain band-pass sees zeros behind the end of the part, not the ringing of the input filter
            if (lane == ns1 && t - ns1 >= len) x = 0.0;
            // scipy _sosfilt: x_cur = b0 x + z0; z0 = b1 x - a1 x_cur + z1; z1 = b2 x - a2 x_cur
            const double yv = b0 * x + z0;
            z0 = (b1 * x - a1 * yv) + z1;
            z1 = b2 * x - a2 * yv;
            out = yv;
            // lane ns1-1 emits y[t - (ns1-1)], lane nl-1 emits f2[128 + t - (nl-1)]: collect 32 steps per lane slot
            const double y_e = shfl64(out, ns1 - 1), f_e = shfl64(out, nl - 1);
            if (lane == i) { ykeep = y_e; fkeep = f_e; }
        }
        const int ny = t0 + lane - (ns1 - 1);
        if (ny >= 0 && ny < len) y[ny] = ykeep;
        if (ns2 > 0) {
            const int nf = t0 + lane - (nl - 1);
            if (nf >= 0 && nf < len + 128) f2[128 + nf] = fkeep;
        }
    }
}

// The same cascade with 32 / (ns_in + ns_td) parts per warp: a group of consecutive lanes per part, section 0 loads its
// samples itself (eight ahead of the recursion), every other section takes the previous lane's output of the previous
// step by one shuffle, and the two emitting sections store their outputs directly -- one double shuffle per step
// instead of four plus the collect.  Same operations on the same values as roe_filter_kernel (bit-equal outputs).
template <typename PCM>
__global__ void __launch_bounds__(128) roe_filter_wave_kernel(const __grid_constant__ RoeDev p, RoeParts pt, int steps /* uniform: longest part + 128 + lanes */,
                                                              const PCM* __restrict__ pcm, double* __restrict__ ybuf, double* __restrict__ tbuf) {
    const unsigned FULL = 0xffffffffu;
    const int ns1 = p.ns_in, ns2 = p.want_td ? p.ns_td : 0, nl = ns1 + ns2;
    const int gpw = 32 / nl;
    const int lane = threadIdx.x & 31;
    const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int grp = lane / nl, s = lane - grp * nl;
    const int part = warp * gpw + grp;
    const bool live = grp < gpw && part < pt.n_parts;
    const int pi = live ? part : 0;
    const int len = live ? pt.len[pi] : 0;
    const PCM* src = pcm + pt.start[pi];
    double* y = ybuf + pt.yo[pi];
    double* f2 = tbuf + pt.yo[pi] + (int64_t)256 * pi;          // len + 256 values: zeros(128), y, zeros(128) filtered
    double b0 = 0, b1 = 0, b2 = 0, a1 = 0, a2 = 0;
    if (s < ns1) { b0 = p.sos_in[s][0]; b1 = p.sos_in[s][1]; b2 = p.sos_in[s][2]; a1 = p.sos_in[s][4]; a2 = p.sos_in[s][5]; }
    else { const int q = s - ns1; b0 = p.sos_td[q][0]; b1 = p.sos_td[q][1]; b2 = p.sos_td[q][2]; a1 = p.sos_td[q][4]; a2 = p.sos_td[q][5]; }
    if (live && ns2 > 0) for (int i = s; i < 128; i += nl) f2[i] = 0.0;   // a filter at rest stays at rest on the leading zeros
    // what this lane stores: the input filter's last section y[m], m < len; the last section of all f2[128 + m], m < len + 128
    const bool emit_y = live && s == ns1 - 1, emit_f = live && ns2 > 0 && s == nl - 1;
    double* outp = emit_y ? y : f2 + 128;
    const unsigned lim = emit_y ? (unsigned)len : (emit_f ? (unsigned)(len + 128) : 0u);
    // the time-domain band-pass (section ns1) sees zeros behind the end of the part, not the ringing of the input filter
    const int zlim = (ns2 > 0 && s == ns1) ? len : 0x7fffffff;
    const bool first = s == 0;
    double z0 = 0.0, z1 = 0.0, out = 0.0;
    // every state starts at rest and a section's input is zero until its first sample arrives, so no step needs gating
    for (int k0 = 0; k0 < steps; k0 += 8) {
        double xs[8];
#pragma unroll
        for (int u = 0; u < 8; u++) xs[u] = (first && k0 + u < len) ? roe_load(src, k0 + u) : 0.0;
        const int mb = k0 - s;                                  // this section's sample at step k0
        double* ob = outp + mb;                                 // dereferenced only where 0 <= mb + u < lim
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const double prev = __shfl_up_sync(FULL, out, 1);   // output of the previous section at the previous step
            double x = first ? xs[u] : prev;
            if (mb + u >= zlim) x = 0.0;
            // scipy _sosfilt: x_cur = b0 x + z0; z0 = b1 x - a1 x_cur + z1; z1 = b2 x - a2 x_cur
            const double yv = b0 * x + z0;
            z0 = (b1 * x - a1 * yv) + z1;
            z1 = b2 * x - a2 * yv;
            out = yv;
            if ((unsigned)(mb + u) < lim) ob[u] = yv;
        }
    }
}

// pairwise (numpy order) sum of f(n), n = 0..255, spread over the 8 lanes of a frame: lane j adds elements 8i + j of
// each 128-block in ascending i, the lanes combine as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), block 0 + block 1
template <typename F>
__device__ __forceinline__ double roe_sum256(F f, int lane, unsigned gmask) {
    double tot[2];
#pragma unroll
    for (int b = 0; b < 2; b++) {
        double r = f(b * 128 + lane);
#pragma unroll
        for (int i = 1; i < 16; i++) r += f(b * 128 + 8 * i + lane);
        r += __hiloint2double(__shfl_xor_sync(gmask, __double2hiint(r), 1), __shfl_xor_sync(gmask, __double2loint(r), 1));
        r += __hiloint2double(__shfl_xor_sync(gmask, __double2hiint(r), 2), __shfl_xor_sync(gmask, __double2loint(r), 2));
        r += __hiloint2double(__shfl_xor_sync(gmask, __double2hiint(r), 4), __shfl_xor_sync(gmask, __double2loint(r), 4));
        tot[b] = r;
    }
    return tot[0] + tot[1];
}

__global__ void __launch_bounds__(ROE_NT, 2) roe_frame_kernel(const __grid_constant__ RoeDev p, RoeParts pt, const double* __restrict__ ybuf,
                                                           const double* __restrict__ tbuf, const cx<double>* __restrict__ twA,
                                                           const cx<double>* __restrict__ tw256, double* __restrict__ mag,
                                                           double* __restrict__ fout) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<double>* s_ex = reinterpret_cast<cx<double>*>(smem_raw);     // [32][kExSize]
    cx<double>* s_twA = s_ex + 32 * kExSize;
    cx<double>* s_tw256 = s_twA + 128;
    double* s_win = reinterpret_cast<double*>(s_tw256 + 130);
    const int part = blockIdx.y, tid = threadIdx.x;
    const int len = pt.len[part];
    const int T = 1 + len / 128;
    const int t0 = blockIdx.x * 32;
    if (t0 >= T) return;
    for (int i = tid; i < 128; i += ROE_NT) s_twA[i] = twA[i];
    for (int i = tid; i < 129; i += ROE_NT) s_tw256[i] = tw256[i];
    for (int i = tid; i < 256; i += ROE_NT) s_win[i] = p.window[i];
    __syncthreads();
    const double* y = ybuf + pt.yo[part];
    const int fr = tid >> 3, lane = tid & 7, t = t0 + fr;
    const unsigned gmask = 0xffu << ((tid & 31) & ~7);
    // sample n of frame t: centred frames over the part with zeros outside (np.pad, and the zeros(hop) of :672)
    auto xs = [&](int n) -> double { const int s = t * 128 - 128 + n; return (s >= 0 && s < len && t < T) ? y[s] : 0.0; };
    cx<double>* ex = s_ex + fr * kExSize;
    rfft256_passA<double>(lane, xs, s_win, s_twA, ex);
    __syncthreads();
    double* mrow = mag + (pt.fo[part] - part + t) * 129;      // frames of earlier parts: fo - part (fo counts frames + 1)
    rfft256_passB<double>(lane, ex, s_tw256, [&](int k, double re, double im) { if (t < T) mrow[k] = hypot(re, im); });
    if (t >= T) return;
    double* o = fout + (pt.fo[part] + t) * ROE_FO;
    if (p.want_td) {
        double kurt = 0.0, crest = 0.0;
        if (t > 0) {
            const double n = 256.0;
            const double mean = roe_sum256([&](int i) { return xs(i); }, lane, gmask) / n;
            const double m2 = roe_sum256([&](int i) { const double d = xs(i) - mean; return d * d; }, lane, gmask) / n;
            const double m4 = roe_sum256([&](int i) { const double d = xs(i) - mean; return (d * d) * (d * d); }, lane, gmask) / n;
            const double ms = roe_sum256([&](int i) { const double v = xs(i); return v * v; }, lane, gmask) / n;
            double pk = 0.0;
            for (int i = lane; i < 256; i += 8) pk = fmax(pk, fabs(xs(i)));
            pk = fmax(pk, __hiloint2double(__shfl_xor_sync(gmask, __double2hiint(pk), 1), __shfl_xor_sync(gmask, __double2loint(pk), 1)));
            pk = fmax(pk, __hiloint2double(__shfl_xor_sync(gmask, __double2hiint(pk), 2), __shfl_xor_sync(gmask, __double2loint(pk), 2)));
            pk = fmax(pk, __hiloint2double(__shfl_xor_sync(gmask, __double2hiint(pk), 4), __shfl_xor_sync(gmask, __double2loint(pk), 4)));
            const double lim = 2.220446049250313e-16 * mean;
            kurt = (m2 <= lim * lim) ? nan("") : m4 / (m2 * m2) - 3.0;     // scipy.stats.kurtosis(fisher=True), biased
            crest = pk / (sqrt(ms) + 1e-12);
        }
        const double* f2 = tbuf + pt.yo[part] + (int64_t)256 * part + (int64_t)t * 128;
        const double en = roe_sum256([&](int i) { const double v = f2[i]; return v * v; }, lane, gmask);
        if (lane == 0) { o[1] = kurt; o[2] = crest; o[4] = en; }
    } else if (lane == 0) { o[1] = 0.0; o[2] = 0.0; o[4] = 0.0; }
}

// scipy.signal._peak_finding_utils._local_maxima_1d on x(0..n-1): calls hit(mid) for every peak (plateau midpoint) in
// ascending order until hit returns false
template <typename X, typename Hit>
__device__ __forceinline__ void roe_local_maxima(X x, int n, Hit hit) {
    int i = 1;
    const int imax = n - 1;
    while (i < imax) {
        if (x(i - 1) < x(i)) {
            int ahead = i + 1;
            while (ahead < imax && x(ahead) == x(i)) ahead++;
            if (x(ahead) < x(i)) {
                if (!hit((i + ahead - 1) / 2)) return;
                i = ahead;
            }
        }
        i++;
    }
}

struct RoeScratch {
    double* a;      // [ROE_LMAX]
    double* b;      // [ROE_LMAX]
    double* red;    // [ROE_NT / 32 + 1]
};

// compute_novelty_spectrum_new on the band [blo, bhi] (thread m = frame m; m == T is the appended zero)
__device__ void roe_novelty(const RoeDev& p, const double* __restrict__ mag, int T, double blo, double bhi, double thr,
                            const RoeScratch& s, double& novk, double& novt) {
    const int m = threadIdx.x, L = T + 1;
    const double f_res = p.fs / (double)p.n_fft;
    const int i1 = (int)(py_floordiv(blo, f_res) + 1.0), i2 = (int)py_floordiv(bhi, f_res);
    double nov = 0.0;
    if (m < T) {
        double r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const double* Y = mag + (int64_t)m * 129;
        const int k0 = max(i1 - 1, 0), k1 = min(i2, 127);
        for (int k = k0; k <= k1; k++) {
            const double ya = (k >= i1 && k <= i2) ? Y[k] : 0.0, yb = (k + 1 >= i1 && k + 1 <= i2) ? Y[k + 1] : 0.0;
            double d = yb - ya;
            if (d <= 0.0) d = 0.0;
            const int j = k & 7;
#pragma unroll
            for (int q = 0; q < 8; q++) if (q == j) r[q] += d;
        }
        nov = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    }
    if (m < L) s.a[m] = nov;
    double mx = (m < L) ? nov : 0.0;
    for (int d = 16; d > 0; d >>= 1) mx = fmax(mx, shfl64(mx, (threadIdx.x & 31) ^ d));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s.red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = 0.0;
    for (int w = 0; w < ROE_NT / 32; w++) mx = fmax(mx, s.red[w]);
    double snr = 0.0;
    if (m < L) {
        // compute_local_average: mean of the wl smallest values of the +-M window, summed in ascending order
        const int lo = max(m - p.M, 0), hi = min(m + p.M + 1, L), wl = p.wl;
        double sm[8];
#pragma unroll
        for (int q = 0; q < 8; q++) sm[q] = INFINITY;
        for (int i = lo; i < hi; i++) {
            double v = s.a[i];
#pragma unroll
            for (int q = 0; q < 8; q++) if (q < wl && v < sm[q]) { const double tq = sm[q]; sm[q] = v; v = tq; }
        }
        const int cnt = min(wl, hi - lo);
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < 8; q++) if (q < cnt) acc = (q == 0) ? sm[0] : acc + sm[q];
        double la = (1.0 / (double)wl) * acc;
        if (la <= 0.0) la = mx / 5.0;
        double v = s.a[m];
        if (v == 0.0) v = 1.0;
        if (la == 0.0) la = 1.0;
        snr = v / la;
        s.b[m] = snr;
    }
    __syncthreads();
    if (m < L) s.a[m] = 0.0;
    __syncthreads();
    // find_peaks(prominence=(None, None)) keeps every local maximum; thread m handles the plateau starting at m
    if (m >= 1 && m < L - 1 && s.b[m - 1] < s.b[m]) {
        int ahead = m + 1;
        while (ahead < L - 1 && s.b[ahead] == s.b[m]) ahead++;
        if (s.b[ahead] < s.b[m]) s.a[(m + ahead - 1) / 2] = 1.0;
    }
    __syncthreads();
    novk = 0.0; novt = 0.0;
    if (m < L) {
        const double mask = s.a[m];
        novt = snr * mask;
        double v = 0.0;
        if (snr > thr) { v = snr; if (snr > thr * 1.5) v = thr * 1.5; }
        novk = v * mask;
    }
    __syncthreads();
}

// find_peaks_in_frequency_range for one frame: frequency of the first of the `max_peaks` lowest local maxima of the
// search band that lies strictly inside (rlo, rhi), else 0
__device__ double roe_peak_in_range(const RoeDev& p, const double* __restrict__ Y, double slo, double shi, double rlo, double rhi) {
    const double fn = p.fs / 2.0;
    int b1 = (int)((slo * 129.0) / fn), b2 = (int)((shi * 129.0) / fn);
    b1 = max(b1, 0); b2 = min(b2, 129);
    double f = 0.0;
    int count = 0;
    if (b2 - b1 >= 3)
        roe_local_maxima([&](int i) { return Y[b1 + i]; }, b2 - b1, [&](int mid) {
            const double fr = ((double)(mid + b1) * fn) / 129.0;
            count++;
            if (rlo < fr && fr < rhi) { f = fr; return false; }
            return count < p.max_peaks;
        });
    return f;
}

__global__ void __launch_bounds__(ROE_NT) roe_part_kernel(const __grid_constant__ RoeDev p, RoeParts pt, const double* __restrict__ mag_all,
                                                          double* __restrict__ fout, double* __restrict__ harm, double* __restrict__ pout) {
    __shared__ double s_a[ROE_LMAX], s_b[ROE_LMAX], s_red[ROE_NT / 32 + 1], s_f[ROE_LMAX];
    __shared__ double s_frain;
    const int part = blockIdx.x, m = threadIdx.x;
    const int T = 1 + pt.len[part] / 128, L = T + 1;
    const double* mag = mag_all + (pt.fo[part] - part) * 129;
    const RoeScratch sc{s_a, s_b, s_red};
    double novk, novt;
    const double blo = p.f_natural, bhi = p.f_natural + 300.0;
    roe_novelty(p, mag, T, blo, bhi, p.rain_thr[0], sc, novk, novt);
    double fpk = 0.0;
    if (m < T) fpk = roe_peak_in_range(p, mag + (int64_t)m * 129, p.search0_lo, p.search0_hi, blo, bhi);
    if (m < T && novk != 0.0 && fpk == 0.0) { novk = 0.0; novt = 0.0; }
    if (m < L) s_f[m] = (m < T) ? fpk : 0.0;
    __syncthreads();
    if (m == 0) {
        // find_nonzero_mean: np.mean of the non-zero peak frequencies in frame order
        int n = 0;
        for (int i = 0; i < T; i++) if (s_f[i] != 0.0) s_b[n++] = s_f[i];
        s_frain = n ? (0.0 + np_pairwise<double>([&](int i) { return s_b[i]; }, 0, n)) / (double)n : 0.0;
    }
    __syncthreads();
    const double frain = s_frain;
    double* o = (m < L) ? fout + (pt.fo[part] + m) * ROE_FO : nullptr;
    if (o) { o[6] = novk; o[7] = novt; }
    const bool natural = p.nat_lo <= frain && frain <= p.nat_hi;
    int mh_set = 0;
    for (int i = 1; i < 6; i++) if (frain * (double)(i + 1) + 300.0 > p.op_hi + 100.0) mh_set = i;
    if (m == 0) {
        double* po = pout + (int64_t)part * ROE_PO;
        po[0] = frain; po[1] = natural ? 1.0 : 0.0; po[2] = (double)mh_set;
    }
    double* hrow = harm + (pt.fo[part] * 5);
    for (int hn = 1; hn < 6; hn++) {
        double nx = 0.0, nt_ = 0.0;
        if (natural) {     // block-uniform
            const double f1 = frain * (double)(hn + 1) - 100.0, f2 = f1 + 300.0;
            double slo = frain * (double)(hn + 1) - 200.0, shi = frain * (double)(hn + 1) + 300.0;
            if (slo < p.op_lo) slo = p.op_lo;
            if (shi > p.op_hi) shi = p.op_hi;
            roe_novelty(p, mag, T, f1, f2, p.rain_thr[hn], sc, nx, nt_);
            if (m < T && nx != 0.0 && roe_peak_in_range(p, mag + (int64_t)m * 129, slo, shi, f1, f2) == 0.0) nx = 0.0;
        }
        if (m < L) hrow[(int64_t)(hn - 1) * L + m] = nx;
    }
}

__global__ void roe_state_kernel(RoeParts pt, const double* __restrict__ pout, int mh_in, int* __restrict__ mh_eff, int* __restrict__ mh_out) {
    if (threadIdx.x || blockIdx.x) return;
    int mh = mh_in;
    for (int q = 0; q < pt.n_parts; q++) {
        const int set = (int)pout[(int64_t)q * ROE_PO + 2];
        if (set) mh = set;
        mh_eff[q] = mh;
    }
    *mh_out = mh;
}

__global__ void __launch_bounds__(ROE_NT) roe_combine_kernel(const __grid_constant__ RoeDev p, RoeParts pt, const int* __restrict__ mh_eff,
                                                             const double* __restrict__ harm, double* __restrict__ fout, double* __restrict__ pout) {
    __shared__ double s_e[ROE_LMAX];
    __shared__ int s_drops, s_peaks;
    const int part = blockIdx.x, m = threadIdx.x;
    const int T = 1 + pt.len[part] / 128, L = T + 1;
    double* o = (m < L) ? fout + (pt.fo[part] + m) * ROE_FO : nullptr;
    if (m == 0) { s_drops = 0; s_peaks = 0; }
    if (m < L) s_e[m] = (m < T) ? o[4] : 0.0;
    __syncthreads();
    if (m < L) {
        const double n0 = o[6];
        double tot = n0;
        if (pout[(int64_t)part * ROE_PO + 1] != 0.0) {
            const double* hrow = harm + pt.fo[part] * 5;
            for (int hn = 1; hn < mh_eff[part]; hn++) tot += (n0 == 0.0) ? 0.0 : hrow[(int64_t)(hn - 1) * L + m];
        }
        if (tot > p.rain_thr_hn) tot = p.rain_thr_hn;
        else if (tot < p.rain_thr_hn) tot = 0.0;
        o[0] = tot;
        if (tot >= 1.0) atomicAdd(&s_drops, 1);
        double diff = 0.0, emin = 0.0;
        if (p.want_td && m < T) {
            const int lo = max(1, m - 30), hi = min(T - 1, m + 31);
            if (lo < hi) { emin = s_e[lo]; for (int i = lo + 1; i < hi; i++) emin = fmin(emin, s_e[i]); }
            if (m >= 2) {
                double last = s_e[m - 1];
                if (s_e[m - 2] < s_e[m - 1]) last = s_e[m - 2];
                if (s_e[m] > last) diff = s_e[m] / (last + 1e-12);
            }
        }
        o[3] = diff; o[5] = emin;
        if (m == T) { o[1] = 0.0; o[2] = 0.0; o[4] = 0.0; }
        if (p.want_td && m < T && o[1] > p.kurtosis_thr && o[2] > p.crest_thr && diff > p.diff_energy_thr) atomicAdd(&s_peaks, 1);
    }
    __syncthreads();
    if (m == 0) { pout[(int64_t)part * ROE_PO + 3] = (double)s_drops; pout[(int64_t)part * ROE_PO + 4] = (double)s_peaks; }
}

__global__ void roe_clip_kernel(const __grid_constant__ RoeDev p, int n_clips, const int* __restrict__ clip_part0, const double* __restrict__ pout,
                                double* __restrict__ cout) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_clips) return;
    int count = 0, peaks = 0, raining = 0;
    for (int q = clip_part0[c]; q < clip_part0[c + 1]; q++) {
        count += (int)pout[(int64_t)q * ROE_PO + 3];
        peaks += (int)pout[(int64_t)q * ROE_PO + 4];
        if (count > p.rain_drop_threshold) raining = 1;
    }
    int mod = count, npk = count;
    if (p.want_td) {
        npk = peaks;
        if (p.handle_fn && !raining && (count > p.rain_drop_max_thr || npk > p.rain_peaks_max_thr)) { raining = 1; mod = max(count, npk); }
        if (p.handle_fp && raining && (npk < p.rain_peaks_min_thr || count < p.rain_drop_threshold)) { raining = 0; mod = 0; }
    }
    double* o = cout + (int64_t)c * APT_ROE_CLIP_F;
    o[0] = raining ? (double)mod : 0.0;   // what rain_detection_algo returns
    o[1] = (double)count; o[2] = (double)npk; o[3] = (double)mod; o[4] = (double)raining;
}

}  // namespace apt
