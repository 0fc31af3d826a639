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
#include "apt_kernels.cuh"

namespace apt {

constexpr int BNE_MAX_SOS = 8;
constexpr int BNE_MAX_S = 8;       // subframes per frame
constexpr int BNE_MAX_W = 64;      // ring buffer length
constexpr int BNE_MAX_BANDS = 8;
constexpr int BNE_SEG = 16;        // frames per filter segment
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
                                                         const int64_t* __restrict__ seg_off, const PCM* __restrict__ pcm,
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
    const int fa = seg * BNE_SEG, fb = min(nfr, fa + BNE_SEG);
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

}  // namespace apt
