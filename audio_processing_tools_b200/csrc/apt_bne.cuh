// apt_bne.cuh -- band noise estimator: the reference's BandNoiseEstimator /
// NoiseFrameDetector (edge/band_noise_estimator.py:107-310, :516-986) driven frame by frame as
// BandNoiseEstimatorProcessor.run does (edge/band_noise_processor.py:82-281), batched over clips.
// float64 configuration (the reference default).  hop == frame_len (the processor requires it), so the
// frames tile the clip and the two IIR filters run as continuous streams over it.
//
//   bne_filter_kernel  one thread per (clip, segment of BNE_SEG frames): streaming HPF -> BPF (scipy sosfilt,
//                      DF2T) from `warm` samples before the segment (exactly seeded with zi * x[0] at the clip
//                      start), subframe energies of both signals in numpy's pairwise order, HPF signal kept
//                      in float64 for the FFT
//   bne_fft_kernel     one CTA per frame: rFFT of the HPF frame -> power / magnitude band sums
//                      (fft_rain_from_power :156-181 inputs, M_band_fft / E_band_fft diagnostics)
//   bne_state_kernel   one thread per clip: detector (FFT jump test, dB-rise subframe mask with hold), ring
//                      buffer with TTL, np.quantile + EMA noise estimate, adaptive q, optional smoothing,
//                      telemetry counters, Wiener-like gain (process_frame :770-986)
#pragma once
#include <type_traits>
#include "apt_kernels.cuh"

namespace apt {

constexpr int BNE_MAX_SOS = 8;
constexpr int BNE_MAX_S = 8;       // subframes per frame
constexpr int BNE_MAX_W = 64;      // ring buffer length
constexpr int BNE_MAX_BANDS = 8;
constexpr int BNE_SEG = 16;        // frames per filter segment (default; APT_BNE_SEG overrides it per call)
constexpr int BNE_FRAME_F = 12;    // per-frame float outputs
constexpr int BNE_STATS = 16;

typedef apt_bne_params_t BneDev;   // the device-side parameter block is the ABI struct itself

template <int NS>
__device__ __forceinline__ double bne_sos(const double (*c)[6], double (*z)[2], int ns, double x) {
    for (int s = 0; s < ns; s++) {
        const double y = c[s][0] * x + z[s][0];                       // scipy _sosfilt: plain mul/add, no FMA
        z[s][0] = c[s][1] * x - c[s][4] * y + z[s][1];
        z[s][1] = c[s][2] * x - c[s][5] * y;
        x = y;
    }
    return x;
}

// numpy pairwise sum of 128 squares held as 8 strided accumulators (n <= 128 branch of pairwise_sum)
struct Sub128 {
    double r[8];
    __device__ __forceinline__ void start(int j, double v) { r[j] = v * v; }
    __device__ __forceinline__ void add(int j, double v) { r[j] += v * v; }
    __device__ __forceinline__ double total() const { return ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7])); }
};

template <typename PCM>
__global__ void __launch_bounds__(128) bne_filter_kernel(const __grid_constant__ BneDev p, int n_clips,
                                                         const int64_t* __restrict__ samp_off, const int64_t* __restrict__ fr_off,
                                                         const int64_t* __restrict__ seg_off, int seg_frames, const PCM* __restrict__ pcm,
                                                         double* __restrict__ xhp, double* __restrict__ subEh, double* __restrict__ subEb) {
    const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= seg_off[n_clips]) return;
    int lo = 0, hi = n_clips;                                         // clip of this segment
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (seg_off[mid] <= gi) lo = mid; else hi = mid; }
    const int c = lo;
    const int seg = (int)(gi - seg_off[c]);
    const int64_t base = samp_off[c];
    const int64_t f0 = fr_off[c];
    const int nfr = (int)(fr_off[c + 1] - f0);
    const int fa = seg * seg_frames, fb = min(nfr, fa + seg_frames);
    const int64_t s_begin = (int64_t)fa * p.N, s_end = (int64_t)fb * p.N;
    int64_t s0 = s_begin - p.warm;
    double zh[BNE_MAX_SOS][2], zb[BNE_MAX_SOS][2];
    const bool exact = s0 <= 0;
    if (exact) s0 = 0;
    const double x0 = exact ? (double)load_sample(pcm, base) : 0.0;   // _need_zi_seed: zi * x[0] (:782-788)
    for (int s = 0; s < BNE_MAX_SOS; s++) {
        zh[s][0] = s < p.ns_h ? p.zi_h[s][0] * x0 : 0.0; zh[s][1] = s < p.ns_h ? p.zi_h[s][1] * x0 : 0.0;
        zb[s][0] = s < p.ns_b ? p.zi_b[s][0] * x0 : 0.0; zb[s][1] = s < p.ns_b ? p.zi_b[s][1] * x0 : 0.0;
    }
    for (int64_t s = s0; s < s_begin; s++) {                           // warm-up: outputs discarded
        const double xh = bne_sos<0>(p.sos_h, zh, p.ns_h, (double)load_sample(pcm, base + s));
        bne_sos<0>(p.sos_b, zb, p.ns_b, xh);
    }
    Sub128 ah, ab;
    const int spf = p.N / p.sub_len;   // == S when subhop == subframe_len
    for (int64_t s = s_begin; s < s_end; s++) {
        const double xh = bne_sos<0>(p.sos_h, zh, p.ns_h, (double)load_sample(pcm, base + s));
        const double xb = bne_sos<0>(p.sos_b, zb, p.ns_b, xh);
        xhp[base + s] = xh;
        const int within = (int)((s - s_begin) % p.sub_len);
        const int j = within & 7;
        if (within < 8) { ah.start(j, xh); ab.start(j, xb); } else { ah.add(j, xh); ab.add(j, xb); }
        if (within == p.sub_len - 1) {
            const int64_t sub = (s - s_begin) / p.sub_len;             // subframe index inside the segment
            const int64_t fr = f0 + fa + sub / spf;
            const int k = (int)(sub % spf);
            subEh[fr * BNE_MAX_S + k] = ah.total();
            subEb[fr * BNE_MAX_S + k] = ab.total();
        }
    }
}

// The same streaming filters as a WAVEFRONT over the cascade: the ns_h + ns_b second-order sections of one segment sit
// on consecutive lanes, lane s runs section s one sample behind lane s - 1 and takes its input from it by a shuffle, so
// a sample costs one section's latency (a multiply, an add, a shuffle) instead of the whole cascade's, and a warp holds
// 32 / (ns_h + ns_b) segments.  Every section performs the operations of bne_sos on the same values in the same order:
// the outputs are bit-equal to bne_filter_kernel's.  The last HPF lane stores the HPF signal and accumulates its
// subframe energies, the last BPF lane those of the band signal; the eight strided partial sums of numpy's pairwise
// order live in registers named by the step (k & 7), a per-lane rotation of numpy's (sample & 7), undone once per subframe.
template <typename PCM>
__global__ void __launch_bounds__(128) bne_filter_wave_kernel(const __grid_constant__ BneDev p, int n_clips, int steps /* uniform: longest run + lanes */,
                                                              const int64_t* __restrict__ samp_off, const int64_t* __restrict__ fr_off,
                                                              const int64_t* __restrict__ seg_off, int seg_frames, const PCM* __restrict__ pcm,
                                                              double* __restrict__ xhp, double* __restrict__ subEh, double* __restrict__ subEb) {
    const unsigned FULL = 0xffffffffu;
    const int G = p.ns_h + p.ns_b;                  // lanes per segment (host: ns_h >= 1, G <= 32)
    const int gpw = 32 / G;                         // segments per warp
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int grp = lane / G, s = lane - grp * G;
    const int64_t nseg = seg_off[n_clips];
    const int64_t gi = warp * gpw + grp;
    const bool live = grp < gpw && gi < nseg;
    int64_t base = 0, f0 = 0, s0 = 0;
    int n_total = 0, n_warm = 0, fa = 0;
    double x0 = 0.0;
    if (live) {
        int lo = 0, hi = n_clips;                   // clip of this segment
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (seg_off[mid] <= gi) lo = mid; else hi = mid; }
        const int seg = (int)(gi - seg_off[lo]);
        base = samp_off[lo];
        f0 = fr_off[lo];
        const int nfr = (int)(fr_off[lo + 1] - f0);
        fa = seg * seg_frames;
        const int fb = min(nfr, fa + seg_frames);
        const int64_t s_begin = (int64_t)fa * p.N, s_end = (int64_t)fb * p.N;
        s0 = s_begin - p.warm;
        const bool exact = s0 <= 0;
        if (exact) s0 = 0;
        x0 = exact ? (double)load_sample(pcm, base) : 0.0;            // _need_zi_seed: zi * x[0] (:782-788)
        n_total = (int)(s_end - s0);
        n_warm = (int)(s_begin - s0);
    }
    const bool is_h = s < p.ns_h;
    const int si = is_h ? s : s - p.ns_h;
    const double c0 = is_h ? p.sos_h[si][0] : p.sos_b[si][0], c1 = is_h ? p.sos_h[si][1] : p.sos_b[si][1];
    const double c2 = is_h ? p.sos_h[si][2] : p.sos_b[si][2], c4 = is_h ? p.sos_h[si][4] : p.sos_b[si][4];
    const double c5 = is_h ? p.sos_h[si][5] : p.sos_b[si][5];
    double z0 = (is_h ? p.zi_h[si][0] : p.zi_b[si][0]) * x0, z1 = (is_h ? p.zi_h[si][1] : p.zi_b[si][1]) * x0;
    const bool out_h = live && s == p.ns_h - 1, out_b = live && s == G - 1;
    double* subE = out_h ? subEh : subEb;
    const int spf = p.N / p.sub_len;
    // samples [n_warm, n_total) of the run are this lane's outputs (none for the sections in between)
    const unsigned n_out = (out_h || out_b) ? (unsigned)(n_total - n_warm) : 0u;
    double R[8];
#pragma unroll
    for (int u = 0; u < 8; u++) R[u] = 0.0;
    int within = 0, sub = 0;                        // position inside the subframe, subframe index inside the segment
    double yprev = 0.0;
    const PCM* src = pcm + base + s0;
    double* dst = xhp + base + s0;
    const bool first = s == 0;
    // gated: the first 16 steps, while the later sections still wait for their first sample (their seeded state must
    // not move); afterwards every step updates -- a section running past its last sample only produces unused values
    auto block = [&](int k0, auto gated) {
        double xs[8];                               // section 0: the block's eight samples, loaded ahead of the recursion
#pragma unroll
        for (int u = 0; u < 8; u++) xs[u] = (first && k0 + u < n_total) ? (double)load_sample(src, k0 + u) : 0.0;
        const int mb = k0 - s;                      // this lane's sample of the run at step k0
        double* db = dst + mb;                      // dereferenced only inside the output range
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const double prev = __shfl_up_sync(FULL, yprev, 1);
            const double xin = first ? xs[u] : prev;
            const double y = c0 * xin + z0;                           // scipy _sosfilt: plain mul/add, no FMA
            if (!decltype(gated)::value || mb + u >= 0) {
                z0 = c1 * xin - c4 * y + z1;
                z1 = c2 * xin - c5 * y;
            }
            yprev = y;
            if ((unsigned)(mb + u - n_warm) < n_out) {
                if (out_h) db[u] = y;
                R[u] += y * y;                      // 0.0 + sq == sq: the first eight samples of a subframe start the sums
                if (++within == p.sub_len) {
                    // numpy's partial sum j (samples = j mod 8 of the subframe) is R[(j + s) & 7]
                    double t[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) { t[j] = R[j]; R[j] = 0.0; }
                    const int d = s & 7;
                    const double r0 = t[d & 7], r1 = t[(1 + d) & 7], r2 = t[(2 + d) & 7], r3 = t[(3 + d) & 7];
                    const double r4 = t[(4 + d) & 7], r5 = t[(5 + d) & 7], r6 = t[(6 + d) & 7], r7 = t[(7 + d) & 7];
                    const int64_t fr = f0 + fa + sub / spf;
                    subE[fr * BNE_MAX_S + (sub % spf)] = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
                    within = 0; sub++;
                }
            }
        }
    };
    block(0, std::true_type{});
    block(8, std::true_type{});
    block(16, std::true_type{});
    block(24, std::true_type{});
    for (int k0 = 32; k0 < steps; k0 += 8) block(k0, std::false_type{});
}

constexpr int BNE_NT = 128;
__global__ void __launch_bounds__(BNE_NT) bne_fft_kernel(const __grid_constant__ BneDev p, const int64_t* __restrict__ samp_off,
                                                         const int64_t* __restrict__ fr_off, const double* __restrict__ xhp,
                                                         const cx<double>* __restrict__ tw, double* __restrict__ fftq /*[nF][4]*/) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = p.N, H = N >> 1;
    cx<double>* bufA = reinterpret_cast<cx<double>*>(smem_raw);
    cx<double>* bufB = bufA + H;
    double* s_P = reinterpret_cast<double*>(bufB + H);   // [H+1] power
    double* s_M = s_P + (H + 1);                         // [H+1] magnitude
    const int tid = threadIdx.x;
    const int c = blockIdx.y;
    const int64_t f0 = fr_off[c];
    const int nfr = (int)(fr_off[c + 1] - f0);
    const int i = blockIdx.x;
    if (i >= nfr) return;
    const double* x = xhp + samp_off[c] + (int64_t)i * N;
    for (int n = tid; n < H; n += BNE_NT) bufA[n] = {x[2 * n], x[2 * n + 1]};
    __syncthreads();
    cx<double>* a = bufA;
    cx<double>* b = bufB;
    for (int q = 1; q < H; q <<= 1) {
        const int tstep = H / q;
        for (int j0 = tid; j0 < (H >> 1); j0 += BNE_NT) {
            const int k = j0 & (q - 1);
            const int j = ((j0 - k) << 1) + k;
            const cx<double> u0 = a[j0];
            const cx<double> xv = a[j0 + (H >> 1)];
            const cx<double> u1 = (k == 0) ? xv : cmul(xv, tw[k * tstep]);
            b[j] = cadd(u0, u1);
            b[j + q] = csub(u0, u1);
        }
        __syncthreads();
        cx<double>* t = a; a = b; b = t;
    }
    for (int k = tid; k <= H; k += BNE_NT) {
        double re, im;
        if (k == 0) { re = a[0].x + a[0].y; im = 0.0; }
        else if (k == H) { re = a[0].x - a[0].y; im = 0.0; }
        else {
            const cx<double> zk = a[k], cn = cconj(a[H - k]);
            const cx<double> e = {(zk.x + cn.x) * 0.5, (zk.y + cn.y) * 0.5};
            const cx<double> d = csub(zk, cn);
            const cx<double> od = {d.y * 0.5, -d.x * 0.5};
            const cx<double> wo = cmul(od, tw[k]);
            re = e.x + wo.x; im = e.y + wo.y;
        }
        s_P[k] = re * re + im * im;
        s_M[k] = hypot(re, im);
    }
    __syncthreads();
    if (tid == 0) {
        auto band = [&](const double* v, int b0, int b1) -> double {
            b0 = max(0, min(b0, H)); b1 = max(0, min(b1, H));
            if (b1 < b0) return 0.0;
            return 0.0 + np_pairwise<double>([&](int k) { return v[k]; }, b0, b1 - b0 + 1);
        };
        double rain = 0.0;
        for (int q = 0; q < p.n_bands; q++) rain += band(s_P, p.band_b0[q], p.band_b1[q]);
        double* o = fftq + (f0 + i) * 4;
        o[0] = rain;
        o[1] = band(s_P, p.prim_b0, p.prim_b1);
        o[2] = p.mask_b1 >= p.mask_b0 ? 0.0 + np_pairwise<double>([&](int k) { return s_M[k]; }, p.mask_b0, p.mask_b1 - p.mask_b0 + 1) : 0.0;
        o[3] = p.mask_b1 >= p.mask_b0 ? 0.0 + np_pairwise<double>([&](int k) { return s_P[k]; }, p.mask_b0, p.mask_b1 - p.mask_b0 + 1) : 0.0;
    }
}

// frame_len = 256 (the sensor's configuration): the 8-lanes-per-frame float64 rFFT of the main path (apt_math.cuh:
// radix-16 pass, swizzled exchange, radix-8 pass + real unpack), 32 frames per CTA; power / magnitude rows replace the
// frame's exchange area, and four lanes of the frame take one band sum each (numpy's pairwise order).  The generic
// kernel above keeps one CTA per frame and serialises the band sums on one thread.
constexpr int BNE_F256_TF = 32;
constexpr int BNE_F256_NT = BNE_F256_TF * 8;
constexpr int BNE_F256_AREA = 130;   // complex elements per frame area: 128 exchange slots; then 2 x 129 doubles
constexpr size_t bne_fft256_smem() { return sizeof(cx<double>) * ((size_t)BNE_F256_TF * BNE_F256_AREA + 128 + 130) + sizeof(double) * 256; }
__global__ void __launch_bounds__(BNE_F256_NT) bne_fft256_kernel(const __grid_constant__ BneDev p, const int64_t* __restrict__ samp_off,
                                                                 const int64_t* __restrict__ fr_off, const double* __restrict__ xhp,
                                                                 const cx<double>* __restrict__ tw, double* __restrict__ fftq /*[nF][4]*/) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<double>* s_ex = reinterpret_cast<cx<double>*>(smem_raw);
    cx<double>* s_tw128 = s_ex + (size_t)BNE_F256_TF * BNE_F256_AREA;   // [k1][lane]: W128^(lane * k1)
    cx<double>* s_tw256 = s_tw128 + 128;
    double* s_one = reinterpret_cast<double*>(s_tw256 + 130);
    const int tid = threadIdx.x;
    const int c = blockIdx.y;
    const int64_t f0 = fr_off[c];
    const int nfr = (int)(fr_off[c + 1] - f0);
    const int t0 = (int)blockIdx.x * BNE_F256_TF;
    if (t0 >= nfr) return;
    for (int i = tid; i < 256; i += BNE_F256_NT) s_one[i] = 1.0;
    for (int i = tid; i < 128; i += BNE_F256_NT) {
        const int k = 2 * (((i & 7) * (i >> 3)) & 127);                  // W128^m = W256^(2m); the table holds k <= 128
        const cx<double> w = tw[k <= 128 ? k : k - 128];
        s_tw128[i] = k <= 128 ? w : cx<double>{-w.x, -w.y};
    }
    for (int i = tid; i < 129; i += BNE_F256_NT) s_tw256[i] = tw[i];
    __syncthreads();
    const int fr = tid >> 3, lane = tid & 7;
    const int i_fr = min(t0 + fr, nfr - 1);                              // slots past the clip end redo its last frame (not stored)
    const double* x = xhp + samp_off[c] + (int64_t)i_fr * 256;
    cx<double>* ex = s_ex + (size_t)fr * BNE_F256_AREA;
    rfft256_passA<double>(lane, [&](int n) { return __ldg(x + n); }, s_one, s_tw128, ex);
    __syncwarp();
    double* s_P = reinterpret_cast<double*>(ex);                          // [129] power, then [129] magnitude
    double* s_M = s_P + 129;
    rfft256_passB<double>(lane, ex, s_tw256,
                          [&](int k, double re, double im) { s_P[k] = re * re + im * im; s_M[k] = hypot(re, im); },
                          [&]() { __syncwarp(); });
    __syncwarp();
    if (t0 + fr >= nfr || lane >= 4) return;
    const int H = 128;
    auto band = [&](const double* v, int b0, int b1) -> double {
        b0 = max(0, min(b0, H)); b1 = max(0, min(b1, H));
        if (b1 < b0) return 0.0;
        return 0.0 + np_pairwise<double>([&](int k) { return v[k]; }, b0, b1 - b0 + 1);
    };
    double r = 0.0;
    if (lane == 0) { for (int q = 0; q < p.n_bands; q++) r += band(s_P, p.band_b0[q], p.band_b1[q]); }
    else if (lane == 1) r = band(s_P, p.prim_b0, p.prim_b1);
    else r = p.mask_b1 >= p.mask_b0 ? 0.0 + np_pairwise<double>([&](int k) { return (lane == 2 ? s_M : s_P)[k]; }, p.mask_b0, p.mask_b1 - p.mask_b0 + 1) : 0.0;
    fftq[(f0 + t0 + fr) * 4 + lane] = r;
}

// frame_len = 512 (the reference's default): 16 lanes per frame, 16 frames per CTA.  The real frame is 256 complex
// points z[m] = x[2m] + i x[2m+1] = a 16 x 16 transform: pass A (lane j) fft16 over q of z[j + 16 q], twiddle
// W256^(j k1), exchange row k1 (column XOR-swizzled by the row: both sides conflict-free); pass B (lane k1) fft16 over j
// -> Z[k1 + 16 k2], written back in natural order; every lane then unpacks its bins k = lane + 16 m (m < 8) together
// with their partners 256 - k.  Power / magnitude rows replace the frame's area; four lanes take one band sum each.
constexpr int BNE_F512_TF = 16;
constexpr int BNE_F512_NT = BNE_F512_TF * 16;
constexpr int BNE_F512_AREA = 258;   // complex elements per frame area: 256 exchange / spectrum slots; then 2 x 257 doubles
constexpr size_t bne_fft512_smem() { return sizeof(cx<double>) * ((size_t)BNE_F512_TF * BNE_F512_AREA + 256 + 258); }
__global__ void __launch_bounds__(BNE_F512_NT) bne_fft512_kernel(const __grid_constant__ BneDev p, const int64_t* __restrict__ samp_off,
                                                                 const int64_t* __restrict__ fr_off, const double* __restrict__ xhp,
                                                                 const cx<double>* __restrict__ tw /* W512^k, k <= 256 */,
                                                                 double* __restrict__ fftq /*[nF][4]*/) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<double>* s_ex = reinterpret_cast<cx<double>*>(smem_raw);
    cx<double>* s_twA = s_ex + (size_t)BNE_F512_TF * BNE_F512_AREA;     // [k1][j]: W256^(j * k1)
    cx<double>* s_tw512 = s_twA + 256;                                   // [257]
    const int tid = threadIdx.x;
    const int c = blockIdx.y;
    const int64_t f0 = fr_off[c];
    const int nfr = (int)(fr_off[c + 1] - f0);
    const int t0 = (int)blockIdx.x * BNE_F512_TF;
    if (t0 >= nfr) return;
    for (int i = tid; i < 256; i += BNE_F512_NT) {
        const int k = 2 * (((i & 15) * (i >> 4)) & 255);                 // W256^m = W512^(2m); the table holds k <= 256
        const cx<double> w = tw[k <= 256 ? k : k - 256];
        s_twA[i] = k <= 256 ? w : cx<double>{-w.x, -w.y};
    }
    for (int i = tid; i < 257; i += BNE_F512_NT) s_tw512[i] = tw[i];
    __syncthreads();
    const int fr = tid >> 4, lane = tid & 15;
    const int i_fr = min(t0 + fr, nfr - 1);                              // slots past the clip end redo its last frame (not stored)
    cx<double>* ex = s_ex + (size_t)fr * BNE_F512_AREA;
    const unsigned fmask = 0xffffu << (threadIdx.x & 16);                // the frame's 16 lanes of this warp
    cx<double> a[16];
    {
        const double* x = xhp + samp_off[c] + (int64_t)i_fr * 512;
#pragma unroll
        for (int q = 0; q < 16; q++) { const int m = lane + 16 * q; a[q].x = __ldg(x + 2 * m); a[q].y = __ldg(x + 2 * m + 1); }
    }
    fft16(a);
#pragma unroll
    for (int k1 = 0; k1 < 16; k1++) {
        const cx<double> v = (k1 == 0 || lane == 0) ? a[k1] : cmul(a[k1], s_twA[k1 * 16 + lane]);
        ex[k1 * 16 + (lane ^ k1)] = v;
    }
    __syncwarp(fmask);
#pragma unroll
    for (int j = 0; j < 16; j++) a[j] = ex[lane * 16 + (j ^ lane)];
    __syncwarp(fmask);
    fft16(a);                                                            // a[k2] = Z[lane + 16 k2]
#pragma unroll
    for (int k2 = 0; k2 < 16; k2++) ex[lane + 16 * k2] = a[k2];
    __syncwarp(fmask);
    // unpack: X[k] = E + W512^k O, X[256 - k] = conj(E - W512^k O), E = (Z[k] + conj Z[256-k]) / 2, O = (Z[k] - conj Z[256-k]) / 2i
    cx<double> zk[9], zn[9];
#pragma unroll
    for (int m = 0; m < 8; m++) { const int k = lane + 16 * m; zk[m] = ex[k]; zn[m] = ex[(256 - k) & 255]; }
    zk[8] = ex[128]; zn[8] = ex[128];
    __syncwarp(fmask);
    double* s_P = reinterpret_cast<double*>(ex);                          // [257] power, then [257] magnitude
    double* s_M = s_P + 257;
    auto put = [&](int k, double re, double im) { s_P[k] = re * re + im * im; s_M[k] = hypot(re, im); };
#pragma unroll
    for (int m = 0; m < 9; m++) {
        if (m == 8 && lane != 0) break;
        const int k = m == 8 ? 128 : lane + 16 * m;
        if (k == 0) { put(0, zk[0].x + zk[0].y, 0.0); put(256, zk[0].x - zk[0].y, 0.0); continue; }
        const cx<double> cn = cconj(zn[m]);
        const cx<double> e = {(zk[m].x + cn.x) * 0.5, (zk[m].y + cn.y) * 0.5};
        const cx<double> d = csub(zk[m], cn);
        const cx<double> o = {d.y * 0.5, -d.x * 0.5};
        const cx<double> wo = cmul(o, s_tw512[k]);
        put(k, e.x + wo.x, e.y + wo.y);
        if (k != 128) put(256 - k, e.x - wo.x, -(e.y - wo.y));
    }
    __syncwarp(fmask);
    if (t0 + fr >= nfr || lane >= 4) return;
    const int H = 256;
    auto band = [&](const double* v, int b0, int b1) -> double {
        b0 = max(0, min(b0, H)); b1 = max(0, min(b1, H));
        if (b1 < b0) return 0.0;
        return 0.0 + np_pairwise<double>([&](int k) { return v[k]; }, b0, b1 - b0 + 1);
    };
    double r = 0.0;
    if (lane == 0) { for (int q = 0; q < p.n_bands; q++) r += band(s_P, p.band_b0[q], p.band_b1[q]); }
    else if (lane == 1) r = band(s_P, p.prim_b0, p.prim_b1);
    else r = p.mask_b1 >= p.mask_b0 ? 0.0 + np_pairwise<double>([&](int k) { return (lane == 2 ? s_M : s_P)[k]; }, p.mask_b0, p.mask_b1 - p.mask_b0 + 1) : 0.0;
    fftq[(f0 + t0 + fr) * 4 + lane] = r;
}

// np.quantile(v[0..n), q), method "linear", on a sorted array (numpy/lib/_function_base_impl.py: virtual index
// n*q + (alpha + q*(1 - alpha - beta)) - 1 with alpha = beta = 1, _lerp with its t >= 0.5 branch)
__device__ __forceinline__ double np_quantile_sorted(const double* v, int n, double q) {
    const double vi = (double)n * q + (1.0 + q * (1.0 - 1.0 - 1.0)) - 1.0;
    double prev = floor(vi);
    double t = vi - prev;
    int ip = (int)prev, in_ = ip + 1;
    if (vi >= (double)(n - 1)) { ip = n - 1; in_ = n - 1; }
    if (vi < 0.0) { ip = 0; in_ = 0; }
    ip = max(0, min(ip, n - 1)); in_ = max(0, min(in_, n - 1));
    const double a = v[ip], b = v[in_];
    const double d = b - a;
    double r = a + d * t;
    if (t >= 0.5) r = b - d * (1.0 - t);
    if (t == 0.0) r = a;   // exact at integer indexes either way
    return r;
}

// Per-frame outputs [nF][BNE_FRAME_F]: M_band, E_band, N_E, N_E_raw, G_mag, M_clean, q_eff, M_band_fft, E_band_fft,
// E_hpf, N_sub, fft_rain (0/1); rain mask bits [nF] (bit s = subframe s); stats [n_clips][BNE_STATS].
__global__ void bne_state_kernel(const __grid_constant__ BneDev p, int n_clips, const int64_t* __restrict__ fr_off,
                                 const double* __restrict__ subEh, const double* __restrict__ subEb, const double* __restrict__ fftq,
                                 double* __restrict__ fo, uint8_t* __restrict__ maskbits, double* __restrict__ stats) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_clips) return;
    const int64_t f0 = fr_off[c];
    const int nfr = (int)(fr_off[c + 1] - f0);
    const int S = p.S, W = p.W;
    const double EPS = 1e-12;
    double buf[BNE_MAX_W];
    long long bidx[BNE_MAX_W];
    bool valid[BNE_MAX_W];
    for (int i = 0; i < W; i++) { buf[i] = 0.0; bidx[i] = -1; valid[i] = false; }
    int wr = 0, count_valid = 0, since = 0, hold = 0;
    double noise_ema = 0.0, q_eff = p.q, ne_smooth = 0.0;
    bool has_prev_fft = false, has_prev_Eb = false, has_prev_L = false;
    double prev_rain = 0.0, prev_prim = 0.0, prev_Eb = 0.0, prev_Lb = 0.0, prev_Lh = 0.0;
    double noise_sum = 0.0, rain_sum = 0.0, total_sum = 0.0;
    long long noise_frames = 0, rain_frames = 0, total_frames = 0, min_valid = 0, underflow = 0, learned_total = 0, repl_total = 0;
    // The valid entries are also kept as a sorted array, updated by one removal / insertion per ring operation: the noise
    // level is np.quantile of the valid entries EVERY frame; sorting 64 values from scratch per frame (an insertion sort
    // in local memory, ~1 000 dependent steps) cost 41.7 -> 34.0 ms on 512 clips x 60 s.
    double sv[BNE_MAX_W];
    for (int i = 0; i < BNE_MAX_W; i++) sv[i] = 0.0;
    int ns = 0;
    auto sv_remove = [&](double v) {
        int i = 0;
        while (i < ns && sv[i] != v) i++;
        if (i == ns) return;                      // (cannot happen: every valid entry is in sv)
        for (; i + 1 < ns; i++) sv[i] = sv[i + 1];
        ns--;
    };
    auto sv_insert = [&](double v) {
        int b = ns - 1;
        while (b >= 0 && sv[b] > v) { sv[b + 1] = sv[b]; b--; }
        sv[b + 1] = v;
        ns++;
    };
    auto expire = [&](long long frame_idx) {
        if (p.ttl <= 0 || count_valid <= 0) return;
        // entries enter in frame order, so the slot about to be overwritten holds the oldest one: if it is valid and
        // young enough, nothing is stale
        if (valid[wr] && (frame_idx - bidx[wr]) <= p.ttl) return;
        int n = 0;
        for (int i = 0; i < W; i++)
            if (valid[i] && (frame_idx - bidx[i]) > p.ttl) { sv_remove(buf[i]); valid[i] = false; buf[i] = 0.0; bidx[i] = -1; n++; }
        count_valid = max(0, count_valid - n);
    };
    auto push = [&](double v, long long frame_idx) {
        if (!valid[wr]) count_valid++; else sv_remove(buf[wr]);
        sv_insert(v);
        buf[wr] = v; valid[wr] = true; bidx[wr] = frame_idx;
        wr = (wr + 1) % W;
    };
    for (int i = 0; i < nfr; i++) {
        const long long frame_idx = i + 1;
        const double* sE = subEb + (f0 + i) * BNE_MAX_S;
        const double* sH = subEh + (f0 + i) * BNE_MAX_S;
        const double* fq = fftq + (f0 + i) * 4;
        // frame energies: np.sum over the frame == pairwise tree over its 128-sample block sums
        auto tree = [&](const double* v) {
            if (S == 4 && p.N == 4 * p.sub_len && p.sub_len == 128) return (v[0] + v[1]) + (v[2] + v[3]);
            double s = 0.0; for (int k = 0; k < S; k++) s += v[k]; return s;
        };
        const double Eb = tree(sE), Ehpf = tree(sH);
        // FFT-domain decision (:156-181)
        bool fft_rain = false;
        if (has_prev_fft) fft_rain = (fq[0] > (prev_rain + EPS) * p.M_ratio) && (fq[1] > (prev_prim + EPS) * p.N_ratio);
        prev_rain = fq[0]; prev_prim = fq[1]; has_prev_fft = true;
        // time-domain subframe mask (:188-274)
        unsigned mask = 0u;
        for (int s = 0; s < S; s++) {
            const double e = sE[s] > EPS ? sE[s] : EPS;
            if (hold > 0) { mask |= 1u << s; hold--; }
            bool trig = false;
            const double eh = sH[s];
            if (eh >= p.min_Ehpf && e >= p.min_Eband) {
                const double Lb = 10.0 * log10(e + EPS), Lh = 10.0 * log10(eh + EPS);
                if (has_prev_L) {
                    const double dLb = Lb - prev_Lb, dLh = Lh - prev_Lh;
                    if (dLb >= p.band_rise_db && (dLb - dLh) >= p.excess_rise_db) trig = true;
                }
                prev_Lb = Lb; prev_Lh = Lh; has_prev_L = true;
            } else {
                has_prev_L = false;
            }
            if (!trig && p.use_dE && has_prev_Eb) {
                const double ehm = sH[s] > EPS ? sH[s] : EPS;
                const double dE = e - prev_Eb > 0.0 ? e - prev_Eb : 0.0;
                if (dE / (ehm + EPS) >= p.dE_thr) trig = true;
            }
            if (!trig && p.use_D && has_prev_Eb && e > (prev_Eb + EPS) * p.D_ratio) trig = true;
            if (trig) { mask |= 1u << s; hold = max(hold, max(0, p.k_subframes - 1)); }
            prev_Eb = e; has_prev_Eb = true;
        }
        if (fft_rain) mask = (1u << S) - 1u;
        expire(frame_idx);
        int learned = 0;
        for (int s = 0; s < S; s++)
            if (p.learn_all || !((mask >> s) & 1u)) { push(sE[s] > p.eps ? sE[s] : p.eps, frame_idx); learned++; }
        int repl = 0;
        if (p.replenish && learned == 0 && (!p.replenish_only_not_full || count_valid < W)) {
            double rs[BNE_MAX_S];
            for (int s = 0; s < S; s++) rs[s] = sE[s];
            for (int a = 1; a < S; a++) { const double v = rs[a]; int b = a - 1; while (b >= 0 && rs[b] > v) { rs[b + 1] = rs[b]; b--; } rs[b + 1] = v; }
            const double qn = np_quantile_sorted(rs, S, p.repl_q);
            push(qn > p.eps ? qn : p.eps, frame_idx);
            repl = 1;
        }
        learned_total += learned; repl_total += repl;
        since = (learned + repl > 0) ? 0 : since + 1;
        if (p.q_adapt) {
            if (repl) q_eff = (1.0 - p.q_repl_alpha) * q_eff + p.q_repl_alpha * p.repl_q;
            if (learned) q_eff = (1.0 - p.q_norm_alpha) * q_eff + p.q_norm_alpha * p.q;
            q_eff = q_eff < 1e-6 ? 1e-6 : (q_eff > 1.0 - 1e-6 ? 1.0 - 1e-6 : q_eff);
        }
        // quantile + EMA noise estimate (:662-680)
        expire(frame_idx);
        double nsub = 0.0;
        if (count_valid < p.W_min) { noise_ema = 0.0; ne_smooth = 0.0; }
        else {
            const double qv = np_quantile_sorted(sv, ns, q_eff);
            noise_ema = (1.0 - p.ema_alpha) * noise_ema + p.ema_alpha * qv;
            nsub = noise_ema;
        }
        const double ne_raw = (double)S * nsub;
        double ne = ne_raw;
        if (p.smooth) {
            const bool raining = fft_rain || mask != 0u;
            const double up = raining ? p.att_wet : p.att_dry;
            const double a = ne_raw > ne_smooth ? up : p.release;
            ne_smooth = (1.0 - a) * ne_smooth + a * ne_raw;
            ne = ne_smooth;
        }
        // telemetry (:715-768): np.sum over the selected subframes (n < 8: sequential)
        double rain_e = 0.0, dry_e = 0.0;
        { double r = -0.0; bool any = false; for (int s = 0; s < S; s++) if ((mask >> s) & 1u) { r += sE[s]; any = true; } if (any) rain_e = 0.0 + r; }
        { double r = -0.0; bool any = false; for (int s = 0; s < S; s++) if (!((mask >> s) & 1u)) { r += sE[s]; any = true; } if (any) dry_e = 0.0 + r; }
        total_sum += Eb > 0.0 ? Eb : 0.0;
        rain_sum += rain_e;
        { const double a = ne > 0.0 ? ne : 0.0, b = dry_e > 0.0 ? dry_e : 0.0; noise_sum += a < b ? a : b; }
        min_valid = total_frames == 0 ? count_valid : (min_valid < count_valid ? min_valid : count_valid);
        total_frames++;
        if (count_valid < p.W_min) underflow++;
        if (mask != 0u) rain_frames++; else noise_frames++;
        // Wiener-like gain (:949-956)
        double num = Eb - p.beta * ne; num = num > 0.0 ? num : 0.0;
        double gp = num / (Eb + p.eps); gp = gp < 0.0 ? 0.0 : (gp > 1.0 ? 1.0 : gp);
        double g = sqrt(gp); g = g < p.gain_floor ? p.gain_floor : (g > 1.0 ? 1.0 : g);
        const double Mb = sqrt(Eb > 0.0 ? Eb : 0.0);
        double* o = fo + (f0 + i) * BNE_FRAME_F;
        o[0] = Mb; o[1] = Eb; o[2] = ne; o[3] = ne_raw; o[4] = g; o[5] = Mb * g; o[6] = q_eff;
        o[7] = fq[2]; o[8] = fq[3]; o[9] = Ehpf; o[10] = nsub; o[11] = fft_rain ? 1.0 : 0.0;
        maskbits[f0 + i] = (uint8_t)mask;
    }
    double* st = stats + (size_t)c * BNE_STATS;
    st[0] = noise_sum; st[1] = rain_sum; st[2] = total_sum; st[3] = (double)noise_frames; st[4] = (double)rain_frames;
    st[5] = (double)total_frames; st[6] = (double)count_valid; st[7] = (double)min_valid; st[8] = (double)underflow;
    st[9] = (double)since; st[10] = (double)learned_total; st[11] = (double)repl_total; st[12] = nfr > 0 ? q_eff : 0.0;
    st[13] = 0.0; st[14] = 0.0; st[15] = 0.0;
}

// ---------------------------------------------------------------------------------------------
// bne_state_warp_kernel: the same state machine, one WARP per clip.  The one-thread-per-clip kernel above keeps the
// ring buffer and its sorted copy in local memory (dynamic indices): every push is ~60 dependent local-memory steps and a
// frame makes up to S of them, so 512 clips x 60 s took 28 ms on 512 threads.  Here lane l owns ring slots l and l + 32
// and elements l and l + 32 of the sorted copy in registers; a removal / insertion is a ballot (position) and one
// shuffle-shift, the quantile two shuffles; the 2 S float64 log10 of the dB-rise test run on 2 S lanes at once; every
// lane carries the scalar state redundantly (uniform control flow), lane 0 stores the frame's outputs.  Arithmetic and
// its order are those of the serial kernel (same values selected, same operations), so the outputs are bit-equal to it
// (tests/test_gpu_parity.py::test_band_noise_warp_kernel_equals_serial).
// ---------------------------------------------------------------------------------------------
template <int ST>    // ST = 4: the default four subframes per frame at compile time (loops and shuffles sized for it); 0: any S <= 8
__global__ void __launch_bounds__(32) bne_state_warp_kernel(const __grid_constant__ BneDev p, int n_clips, const int64_t* __restrict__ fr_off,
                                                            const double* __restrict__ subEh, const double* __restrict__ subEb,
                                                            const double* __restrict__ fftq, double* __restrict__ fo,
                                                            uint8_t* __restrict__ maskbits, double* __restrict__ stats) {
    const int c = blockIdx.x;
    if (c >= n_clips) return;
    const unsigned FULL = 0xffffffffu;
    const int l = threadIdx.x;
    const int64_t f0 = fr_off[c];
    const int nfr = (int)(fr_off[c + 1] - f0);
    constexpr int SM = ST ? ST : BNE_MAX_S;
    const int S = ST ? ST : p.S, W = p.W;
    const double EPS = 1e-12;
    // ring slots l (A) and l + 32 (B): value and frame index (-1 = empty); sorted copy elements l (A) and l + 32 (B)
    double bufA = 0.0, bufB = 0.0, svA = 0.0, svB = 0.0;
    int bidxA = -1, bidxB = -1;
    int ns = 0, wr = 0, count_valid = 0, since = 0, hold = 0;
    double noise_ema = 0.0, q_eff = p.q, ne_smooth = 0.0;
    bool has_prev_fft = false, has_prev_Eb = false, has_prev_L = false;
    double prev_rain = 0.0, prev_prim = 0.0, prev_Eb = 0.0, prev_Lb = 0.0, prev_Lh = 0.0;
    double noise_sum = 0.0, rain_sum = 0.0, total_sum = 0.0;
    long long noise_frames = 0, rain_frames = 0, total_frames = 0, min_valid = 0, underflow = 0, learned_total = 0, repl_total = 0;

    auto sv_get = [&](int i) { return __shfl_sync(FULL, (i >> 5) ? svB : svA, i & 31); };
    auto sv_remove = [&](double v) {
        const unsigned mA = __ballot_sync(FULL, l < ns && svA == v);
        const unsigned mB = __ballot_sync(FULL, l + 32 < ns && svB == v);
        if (!(mA | mB)) return;                   // (cannot happen: every valid entry is in the sorted copy)
        const int idx = mA ? __ffs(mA) - 1 : 32 + __ffs(mB) - 1;
        double a_dn = __shfl_down_sync(FULL, svA, 1);
        const double b0 = __shfl_sync(FULL, svB, 0);
        if (l == 31) a_dn = b0;
        const double b_dn = __shfl_down_sync(FULL, svB, 1);
        if (idx < 32) { svA = l >= idx ? a_dn : svA; svB = b_dn; }
        else svB = l >= idx - 32 ? b_dn : svB;
        ns--;
    };
    auto sv_insert = [&](double v) {
        const int pos = __popc(__ballot_sync(FULL, l < ns && svA <= v)) + __popc(__ballot_sync(FULL, l + 32 < ns && svB <= v));
        const double a_up = __shfl_up_sync(FULL, svA, 1);
        double b_up = __shfl_up_sync(FULL, svB, 1);
        const double carry = __shfl_sync(FULL, svA, 31);
        if (l == 0) b_up = carry;
        if (pos < 32) { svA = l < pos ? svA : (l == pos ? v : a_up); svB = b_up; }
        else { const int q = pos - 32; svB = l < q ? svB : (l == q ? v : b_up); }
        ns++;
    };
    auto expire = [&](int frame_idx) {
        if (p.ttl <= 0 || count_valid <= 0) return;
        // entries enter in frame order, so the slot about to be overwritten holds the oldest one
        const int bw = __shfl_sync(FULL, (wr >> 5) ? bidxB : bidxA, wr & 31);
        if (bw >= 0 && (long long)frame_idx - bw <= p.ttl) return;
        int n = 0;
        const bool stA = bidxA >= 0 && (long long)frame_idx - bidxA > p.ttl;
        const bool stB = bidxB >= 0 && (long long)frame_idx - bidxB > p.ttl;
        for (unsigned m = __ballot_sync(FULL, stA); m; m &= m - 1) { sv_remove(__shfl_sync(FULL, bufA, __ffs(m) - 1)); n++; }
        for (unsigned m = __ballot_sync(FULL, stB); m; m &= m - 1) { sv_remove(__shfl_sync(FULL, bufB, __ffs(m) - 1)); n++; }
        if (stA) { bufA = 0.0; bidxA = -1; }
        if (stB) { bufB = 0.0; bidxB = -1; }
        count_valid = max(0, count_valid - n);
    };
    auto push = [&](double v, int frame_idx) {
        const int half = wr >> 5, own = wr & 31;
        const int bw = __shfl_sync(FULL, half ? bidxB : bidxA, own);
        const double old = __shfl_sync(FULL, half ? bufB : bufA, own);
        if (bw < 0) count_valid++; else sv_remove(old);
        sv_insert(v);
        if (l == own) { if (half) { bufB = v; bidxB = frame_idx; } else { bufA = v; bidxA = frame_idx; } }
        wr = (wr + 1) % W;
    };
    auto quantile = [&](double q) {                // np_quantile_sorted on the distributed sorted copy
        const int n = ns;
        const double vi = (double)n * q + (1.0 + q * (1.0 - 1.0 - 1.0)) - 1.0;
        const double prev = floor(vi);
        const double t = vi - prev;
        int ip = (int)prev, in_ = ip + 1;
        if (vi >= (double)(n - 1)) { ip = n - 1; in_ = n - 1; }
        if (vi < 0.0) { ip = 0; in_ = 0; }
        ip = max(0, min(ip, n - 1)); in_ = max(0, min(in_, n - 1));
        const double a = sv_get(ip), b = sv_get(in_);
        const double d = b - a;
        double r = a + d * t;
        if (t >= 0.5) r = b - d * (1.0 - t);
        if (t == 0.0) r = a;
        return r;
    };

    // lane l < 8: band subframe energy l; 8 <= l < 16: HPF subframe energy l - 8; 16 <= l < 20: FFT quantity l - 16.
    // One frame ahead in registers.
    auto load_frame = [&](int i) -> double {
        if (i >= nfr) return 0.0;
        const int64_t f = f0 + i;
        if (l < 8) return l < S ? subEb[f * BNE_MAX_S + l] : 0.0;
        if (l < 16) return l - 8 < S ? subEh[f * BNE_MAX_S + (l - 8)] : 0.0;
        if (l < 20) return fftq[f * 4 + (l - 16)];
        return 0.0;
    };
    const bool fast_tree = S == 4 && p.N == 4 * p.sub_len && p.sub_len == 128;
    double nxt = load_frame(0);
    for (int i = 0; i < nfr; i++) {
        const int frame_idx = i + 1;
        const double cur = nxt;
        nxt = load_frame(i + 1);
        // dB levels of the rise test, 2 S lanes at once (the serial kernel evaluates them only where the energy floors
        // pass; the values are the same wherever it does)
        const double lv = l < 8 ? (cur > EPS ? cur : EPS) : cur;
        const double Ldb = l < 16 ? 10.0 * log10(lv + EPS) : 0.0;
        double sE[SM], sH[SM], LB[SM], LH[SM];
#pragma unroll
        for (int s = 0; s < SM; s++) {
            sE[s] = __shfl_sync(FULL, cur, s); sH[s] = __shfl_sync(FULL, cur, 8 + s);
            LB[s] = __shfl_sync(FULL, Ldb, s); LH[s] = __shfl_sync(FULL, Ldb, 8 + s);
        }
        const double fq0 = __shfl_sync(FULL, cur, 16), fq1 = __shfl_sync(FULL, cur, 17);
        const double fq2 = __shfl_sync(FULL, cur, 18), fq3 = __shfl_sync(FULL, cur, 19);
        double Eb, Ehpf;
        if (fast_tree) { Eb = (sE[0] + sE[1]) + (sE[2] + sE[3]); Ehpf = (sH[0] + sH[1]) + (sH[2] + sH[3]); }
        else {
            Eb = 0.0; Ehpf = 0.0;
#pragma unroll
            for (int k = 0; k < SM; k++) if (k < S) { Eb += sE[k]; Ehpf += sH[k]; }
        }
        bool fft_rain = false;
        if (has_prev_fft) fft_rain = (fq0 > (prev_rain + EPS) * p.M_ratio) && (fq1 > (prev_prim + EPS) * p.N_ratio);
        prev_rain = fq0; prev_prim = fq1; has_prev_fft = true;
        unsigned mask = 0u;
#pragma unroll
        for (int s = 0; s < SM; s++) {
            if (s < S) {
                const double e = sE[s] > EPS ? sE[s] : EPS;
                if (hold > 0) { mask |= 1u << s; hold--; }
                bool trig = false;
                const double eh = sH[s];
                if (eh >= p.min_Ehpf && e >= p.min_Eband) {
                    const double Lb = LB[s], Lh = LH[s];
                    if (has_prev_L) {
                        const double dLb = Lb - prev_Lb, dLh = Lh - prev_Lh;
                        if (dLb >= p.band_rise_db && (dLb - dLh) >= p.excess_rise_db) trig = true;
                    }
                    prev_Lb = Lb; prev_Lh = Lh; has_prev_L = true;
                } else {
                    has_prev_L = false;
                }
                if (!trig && p.use_dE && has_prev_Eb) {
                    const double ehm = sH[s] > EPS ? sH[s] : EPS;
                    const double dE = e - prev_Eb > 0.0 ? e - prev_Eb : 0.0;
                    if (dE / (ehm + EPS) >= p.dE_thr) trig = true;
                }
                if (!trig && p.use_D && has_prev_Eb && e > (prev_Eb + EPS) * p.D_ratio) trig = true;
                if (trig) { mask |= 1u << s; hold = max(hold, max(0, p.k_subframes - 1)); }
                prev_Eb = e; has_prev_Eb = true;
            }
        }
        if (fft_rain) mask = (1u << S) - 1u;
        expire(frame_idx);
        int learned = 0;
#pragma unroll
        for (int s = 0; s < SM; s++)
            if (s < S && (p.learn_all || !((mask >> s) & 1u))) { push(sE[s] > p.eps ? sE[s] : p.eps, frame_idx); learned++; }
        int repl = 0;
        if (p.replenish && learned == 0 && (!p.replenish_only_not_full || count_valid < W)) {
            // np.quantile of the frame's S energies: ranks by counting (ties by index), then the serial kernel's lerp
            const double vi = (double)S * p.repl_q + (1.0 + p.repl_q * (1.0 - 1.0 - 1.0)) - 1.0;
            const double prev = floor(vi);
            const double t = vi - prev;
            int ip = (int)prev, in_ = ip + 1;
            if (vi >= (double)(S - 1)) { ip = S - 1; in_ = S - 1; }
            if (vi < 0.0) { ip = 0; in_ = 0; }
            ip = max(0, min(ip, S - 1)); in_ = max(0, min(in_, S - 1));
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int x = 0; x < SM; x++) {
                if (x < S) {
                    int rank = 0;
#pragma unroll
                    for (int y = 0; y < SM; y++)
                        if (y < S && (sE[y] < sE[x] || (sE[y] == sE[x] && y < x))) rank++;
                    if (rank == ip) a = sE[x];
                    if (rank == in_) b = sE[x];
                }
            }
            const double d = b - a;
            double qn = a + d * t;
            if (t >= 0.5) qn = b - d * (1.0 - t);
            if (t == 0.0) qn = a;
            push(qn > p.eps ? qn : p.eps, frame_idx);
            repl = 1;
        }
        learned_total += learned; repl_total += repl;
        since = (learned + repl > 0) ? 0 : since + 1;
        if (p.q_adapt) {
            if (repl) q_eff = (1.0 - p.q_repl_alpha) * q_eff + p.q_repl_alpha * p.repl_q;
            if (learned) q_eff = (1.0 - p.q_norm_alpha) * q_eff + p.q_norm_alpha * p.q;
            q_eff = q_eff < 1e-6 ? 1e-6 : (q_eff > 1.0 - 1e-6 ? 1.0 - 1e-6 : q_eff);
        }
        expire(frame_idx);
        double nsub = 0.0;
        if (count_valid < p.W_min) { noise_ema = 0.0; ne_smooth = 0.0; }
        else {
            const double qv = quantile(q_eff);
            noise_ema = (1.0 - p.ema_alpha) * noise_ema + p.ema_alpha * qv;
            nsub = noise_ema;
        }
        const double ne_raw = (double)S * nsub;
        double ne = ne_raw;
        if (p.smooth) {
            const bool raining = fft_rain || mask != 0u;
            const double up = raining ? p.att_wet : p.att_dry;
            const double a = ne_raw > ne_smooth ? up : p.release;
            ne_smooth = (1.0 - a) * ne_smooth + a * ne_raw;
            ne = ne_smooth;
        }
        double rain_e = 0.0, dry_e = 0.0;
        {
            double r = -0.0, d = -0.0; bool anyr = false, anyd = false;
#pragma unroll
            for (int s = 0; s < SM; s++) {
                if (s < S) {
                    if ((mask >> s) & 1u) { r += sE[s]; anyr = true; } else { d += sE[s]; anyd = true; }
                }
            }
            if (anyr) rain_e = 0.0 + r;
            if (anyd) dry_e = 0.0 + d;
        }
        total_sum += Eb > 0.0 ? Eb : 0.0;
        rain_sum += rain_e;
        { const double a = ne > 0.0 ? ne : 0.0, b = dry_e > 0.0 ? dry_e : 0.0; noise_sum += a < b ? a : b; }
        min_valid = total_frames == 0 ? count_valid : (min_valid < count_valid ? min_valid : count_valid);
        total_frames++;
        if (count_valid < p.W_min) underflow++;
        if (mask != 0u) rain_frames++; else noise_frames++;
        double num = Eb - p.beta * ne; num = num > 0.0 ? num : 0.0;
        double gp = num / (Eb + p.eps); gp = gp < 0.0 ? 0.0 : (gp > 1.0 ? 1.0 : gp);
        double g = sqrt(gp); g = g < p.gain_floor ? p.gain_floor : (g > 1.0 ? 1.0 : g);
        const double Mb = sqrt(Eb > 0.0 ? Eb : 0.0);
        if (l == 0) {
            double* o = fo + (f0 + i) * BNE_FRAME_F;
            o[0] = Mb; o[1] = Eb; o[2] = ne; o[3] = ne_raw; o[4] = g; o[5] = Mb * g; o[6] = q_eff;
            o[7] = fq2; o[8] = fq3; o[9] = Ehpf; o[10] = nsub; o[11] = fft_rain ? 1.0 : 0.0;
            maskbits[f0 + i] = (uint8_t)mask;
        }
    }
    if (l == 0) {
        double* st = stats + (size_t)c * BNE_STATS;
        st[0] = noise_sum; st[1] = rain_sum; st[2] = total_sum; st[3] = (double)noise_frames; st[4] = (double)rain_frames;
        st[5] = (double)total_frames; st[6] = (double)count_valid; st[7] = (double)min_valid; st[8] = (double)underflow;
        st[9] = (double)since; st[10] = (double)learned_total; st[11] = (double)repl_total; st[12] = nfr > 0 ? q_eff : 0.0;
        st[13] = 0.0; st[14] = 0.0; st[15] = 0.0;
    }
}

}  // namespace apt
