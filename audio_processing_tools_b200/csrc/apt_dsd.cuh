// apt_dsd.cuh -- drop-size-distribution emulator: the compute path of the reference's
// transform.process_audio_file_dsd (transform.py:251-313), i.e. DsdProcessingEmualtor
// (host_analysis/device_dsd_processing_emulator.py:16-314), batched over clips.
//
//   dsd_frame_kernel    one CTA per hop position: (optionally Hann-windowed) frame -> float64 FFT ->
//                       |X[k]| -> drop energy (sum over the rain band), peak bin and its magnitude
//                       (process_audio_frame, :128-180, the spectral part)
//   dsd_fft512_kernel   the same for frame_length = 512 (the emulator's configuration): 16 lanes per frame, 16 frames
//                       per CTA, two radix-16 passes through a swizzled exchange (the band noise estimator's layout)
//   dsd_times_kernel    one thread per hop position: everything of the state machine that depends only on the
//                       position -- the frame's timestamp (Python-float expression order), its 2-second slot, whether
//                       the next frame opens a new slot, its drop-size bin -- so that the serial kernel below has no
//                       fmod / log / divide per frame
//   dsd_minutes_kernel  one thread per clip: the per-minute state machine (rain check windows, frame
//                       skipping, 2-second peak-frequency slots, log-binned drop histogram, FFT energies)
//                       over the per-frame quantities (process_audio_data :257-314 and its helpers)
#pragma once
#include "apt_kernels.cuh"

namespace apt {

struct DsdDev {
    int fs, L, hop, window;
    int n_bins, rain_lo, rain_hi, pft_lo, pft_hi, lwin0, lwin1, hwin0, hwin1;
    double rain_thr, rain_log_factor, rain_log_base;
    int max_minutes;
};

constexpr int DSD_NT = 128;
constexpr int DSD_OUT = 100;   // 32 drop-size bins + 30 peak-frequency slots + 38 FFT energies

__global__ void __launch_bounds__(DSD_NT) dsd_frame_kernel(const __grid_constant__ DsdDev p, int clip0,
                                                           const int64_t* __restrict__ samp_off, const int64_t* __restrict__ fr_off,
                                                           const int16_t* __restrict__ pcm, const double* __restrict__ win,
                                                           const cx<double>* __restrict__ tw,
                                                           double* __restrict__ drop, int* __restrict__ pk_idx, double* __restrict__ pk_val) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = p.L, H = N >> 1;
    cx<double>* bufA = reinterpret_cast<cx<double>*>(smem_raw);
    cx<double>* bufB = bufA + H;
    double* s_mag = reinterpret_cast<double*>(bufB + H);   // [H] magnitudes of bins 0..H-1
    const int tid = threadIdx.x;
    const int c = clip0 + (int)blockIdx.y;
    const int64_t base = samp_off[c];
    const int64_t f0 = fr_off[c];
    const int nfr = (int)(fr_off[c + 1] - f0);
    const int i = blockIdx.x;
    if (i >= nfr) return;
    const int64_t s0 = base + (int64_t)i * p.hop;
    // parse.pcm_to_float: int16 / 32768 in float64 (exact)
    for (int n = tid; n < H; n += DSD_NT) {
        double xa = (double)pcm[s0 + 2 * n] * (1.0 / 32768.0), xb = (double)pcm[s0 + 2 * n + 1] * (1.0 / 32768.0);
        if (p.window) { xa *= win[2 * n]; xb *= win[2 * n + 1]; }
        bufA[n] = {xa, xb};
    }
    __syncthreads();
    cx<double>* x = bufA;
    cx<double>* y = bufB;
    for (int q = 1; q < H; q <<= 1) {
        const int tstep = H / q;
        for (int j0 = tid; j0 < (H >> 1); j0 += DSD_NT) {
            const int k = j0 & (q - 1);
            const int j = ((j0 - k) << 1) + k;
            const cx<double> u0 = x[j0];
            const cx<double> xv = x[j0 + (H >> 1)];
            const cx<double> u1 = (k == 0) ? xv : cmul(xv, tw[k * tstep]);
            y[j] = cadd(u0, u1);
            y[j + q] = csub(u0, u1);
        }
        __syncthreads();
        cx<double>* t = x; x = y; y = t;
    }
    // |X[k]| for the bins the emulator looks at (np.abs of complex128 = hypot)
    const int klo = min(p.pft_lo, p.rain_lo), khi = max(p.pft_hi - 1, p.rain_hi);
    for (int k = klo + tid; k <= khi && k < H; k += DSD_NT) {
        double re, im;
        if (k == 0) { re = x[0].x + x[0].y; im = 0.0; }
        else {
            const cx<double> zk = x[k], cn = cconj(x[H - k]);
            const cx<double> e = {(zk.x + cn.x) * 0.5, (zk.y + cn.y) * 0.5};
            const cx<double> d = csub(zk, cn);
            const cx<double> od = {d.y * 0.5, -d.x * 0.5};
            const cx<double> wo = cmul(od, tw[k]);
            re = e.x + wo.x; im = e.y + wo.y;
        }
        s_mag[k] = hypot(re, im);
    }
    __syncthreads();
    if (tid == 0) {
        double de = 0.0;
        for (int k = p.rain_lo; k <= p.rain_hi; k++) de += s_mag[k];   // sequential, like the reference loop
        int best = p.pft_lo;
        for (int k = p.pft_lo + 1; k < p.pft_hi; k++) if (s_mag[k] > s_mag[best]) best = k;   // np.argmax: first maximum
        drop[f0 + i] = de; pk_idx[f0 + i] = best; pk_val[f0 + i] = s_mag[best];
    }
}

// frame_length = 512: 16 lanes per frame (see bne_fft512_kernel for the transform), 16 frames per CTA; the magnitudes
// of the bins the emulator reads replace the frame's exchange area, lane 0 of the frame does the two sequential scans.
constexpr int DSD_F512_TF = 16;
constexpr int DSD_F512_NT = DSD_F512_TF * 16;
constexpr int DSD_F512_AREA = 258;
constexpr size_t dsd_fft512_smem() { return sizeof(cx<double>) * ((size_t)DSD_F512_TF * DSD_F512_AREA + 256 + 258) + sizeof(double) * 512; }
__global__ void __launch_bounds__(DSD_F512_NT) dsd_fft512_kernel(const __grid_constant__ DsdDev p, const int64_t* __restrict__ samp_off,
                                                                 const int64_t* __restrict__ fr_off, const int16_t* __restrict__ pcm,
                                                                 const double* __restrict__ win, const cx<double>* __restrict__ tw,
                                                                 double* __restrict__ drop, int* __restrict__ pk_idx, double* __restrict__ pk_val) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<double>* s_ex = reinterpret_cast<cx<double>*>(smem_raw);
    cx<double>* s_twA = s_ex + (size_t)DSD_F512_TF * DSD_F512_AREA;     // [k1][j]: W256^(j * k1)
    cx<double>* s_tw512 = s_twA + 256;                                   // [257]
    double* s_win = reinterpret_cast<double*>(s_tw512 + 258);            // [512]
    const int tid = threadIdx.x;
    const int c = blockIdx.y;
    const int64_t f0 = fr_off[c];
    const int nfr = (int)(fr_off[c + 1] - f0);
    const int t0 = (int)blockIdx.x * DSD_F512_TF;
    if (t0 >= nfr) return;
    for (int i = tid; i < 256; i += DSD_F512_NT) {
        const int k = 2 * (((i & 15) * (i >> 4)) & 255);                 // W256^m = W512^(2m); the table holds k <= 256
        const cx<double> w = tw[k <= 256 ? k : k - 256];
        s_twA[i] = k <= 256 ? w : cx<double>{-w.x, -w.y};
    }
    for (int i = tid; i < 257; i += DSD_F512_NT) s_tw512[i] = tw[i];
    if (p.window) for (int i = tid; i < 512; i += DSD_F512_NT) s_win[i] = win[i];
    __syncthreads();
    const int fr = tid >> 4, lane = tid & 15;
    const int i_fr = min(t0 + fr, nfr - 1);                              // slots past the clip end redo its last frame (not stored)
    const int16_t* x = pcm + samp_off[c] + (int64_t)i_fr * p.hop;
    cx<double>* ex = s_ex + (size_t)fr * DSD_F512_AREA;
    const unsigned fmask = 0xffffu << (threadIdx.x & 16);                // the frame's 16 lanes of this warp
    cx<double> a[16];
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const int m = lane + 16 * q;
        // parse.pcm_to_float: int16 / 32768 in float64 (exact)
        double xa = (double)__ldg(x + 2 * m) * (1.0 / 32768.0), xb = (double)__ldg(x + 2 * m + 1) * (1.0 / 32768.0);
        if (p.window) { xa *= s_win[2 * m]; xb *= s_win[2 * m + 1]; }
        a[q] = {xa, xb};
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
    cx<double> zk[8], zn[8];
#pragma unroll
    for (int m = 0; m < 8; m++) { const int k = lane + 16 * m; zk[m] = ex[k]; zn[m] = ex[(256 - k) & 255]; }
    const cx<double> z128 = ex[128];
    __syncwarp(fmask);
    double* s_mag = reinterpret_cast<double*>(ex);                        // [256] magnitudes of bins 0..255 (np.abs = hypot)
    // only the bins the emulator reads
    const int klo = min(p.pft_lo, p.rain_lo), khi = max(p.pft_hi - 1, p.rain_hi);
#pragma unroll
    for (int m = 0; m < 8; m++) {
        const int k = lane + 16 * m;                                     // bins k and 256 - k
        const bool want_k = k >= klo && k <= khi, want_n = k > 0 && 256 - k >= klo && 256 - k <= khi;
        if (!want_k && !want_n) continue;
        if (k == 0) { s_mag[0] = hypot(zk[0].x + zk[0].y, 0.0); continue; }
        const cx<double> cn = cconj(zn[m]);
        const cx<double> e = {(zk[m].x + cn.x) * 0.5, (zk[m].y + cn.y) * 0.5};
        const cx<double> d = csub(zk[m], cn);
        const cx<double> o = {d.y * 0.5, -d.x * 0.5};
        const cx<double> wo = cmul(o, s_tw512[k]);
        if (want_k) s_mag[k] = hypot(e.x + wo.x, e.y + wo.y);
        if (want_n) s_mag[256 - k] = hypot(e.x - wo.x, -(e.y - wo.y));
    }
    if (lane == 0 && klo <= 128 && khi >= 128) {
        // bin 128 pairs with itself: E = Re Z, O = Im Z
        const cx<double> wo = cmul(cx<double>{z128.y, -0.0}, s_tw512[128]);
        s_mag[128] = hypot(z128.x + wo.x, 0.0 + wo.y);
    }
    __syncwarp(fmask);
    if (t0 + fr >= nfr || lane != 0) return;
    double de = 0.0;
    for (int k = p.rain_lo; k <= p.rain_hi; k++) de += s_mag[k];         // sequential, like the reference loop
    int best = p.pft_lo;
    for (int k = p.pft_lo + 1; k < p.pft_hi; k++) if (s_mag[k] > s_mag[best]) best = k;   // np.argmax: first maximum
    drop[f0 + t0 + fr] = de; pk_idx[f0 + t0 + fr] = best; pk_val[f0 + t0 + fr] = s_mag[best];
}

// Position-only quantities of the state machine, one thread per hop position (and one slot past the clip's last frame
// for the timestamp the machine computes after it).  ts_cur of position 0 is the clip's timestamp itself, later
// positions use the reference's expression ts_start + (frame_count * hop) / fs (:300-306) -- both in Python-float order.
// aux: bits 0..7 the frame's 2-second slot, bit 8 "the next frame opens another slot", bits 16..23 drop-size bin + 1 (0: no drop).
__global__ void __launch_bounds__(256) dsd_times_kernel(const __grid_constant__ DsdDev p, const int64_t* __restrict__ fr_off,
                                                        const double* __restrict__ ts_in, const double* __restrict__ drop,
                                                        double* __restrict__ tsc /*[nF + n_clips]*/, int* __restrict__ aux /*[nF]*/) {
    const int c = blockIdx.y;
    const int64_t f0 = fr_off[c];
    const int nfr = (int)(fr_off[c + 1] - f0);
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos > nfr) return;
    const double fs = (double)p.fs, hop = (double)p.hop;
    const double hop_s = hop / fs;
    const double ts = ts_in[c];
    const double ts_start = ts - fmod(ts, 60.0);
    const long long fc0 = (long long)(fmod(ts, 60.0) * fs / hop);
    const double ts_cur = pos == 0 ? ts : ts_start + (double)((fc0 + pos) * p.hop) / fs;
    tsc[f0 + c + pos] = ts_cur;
    if (pos == nfr) return;
    const int nxt = (int)(fmod(ts_cur + hop_s, 60.0) / 2.0);
    const int cur = (int)(fmod(ts_cur, 60.0) / 2.0);
    int h1 = 0;
    const double de = drop[f0 + pos];
    if (de > p.rain_thr) {
        int h = (int)floor(log(1.0 + (de - p.rain_thr) * p.rain_log_factor) / log(p.rain_log_base));
        h = h > 31 ? 31 : (h < 0 ? 0 : h);
        h1 = h + 1;
    }
    aux[f0 + pos] = (cur & 0xff) | (nxt != cur ? 0x100 : 0) | (h1 << 16);
}

// One thread per clip.  Every floating-point expression is evaluated in float64 in the reference's order
// (Python floats), because the comparisons against minute / rain-check boundaries depend on the rounding.
__global__ void dsd_minutes_kernel(const __grid_constant__ DsdDev p, int n_clips, const int64_t* __restrict__ samp_off,
                                   const int64_t* __restrict__ fr_off, const double* __restrict__ ts_in,
                                   const double* __restrict__ tsc, const int* __restrict__ aux, const int* __restrict__ pk_idx,
                                   const double* __restrict__ pk_val, double* __restrict__ out, int* __restrict__ n_minutes) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_clips) return;
    const int64_t n = samp_off[c + 1] - samp_off[c];
    const int64_t f0 = fr_off[c];
    double* o = out + (size_t)c * p.max_minutes * DSD_OUT;
    int produced = 0;
    if (n < p.L) { n_minutes[c] = 0; return; }
    unsigned short peak_hist[2048];
    double freq_hist[2048];
    double energy[DSD_OUT];
    const double fs = (double)p.fs, hop = (double)p.hop;
    const double hop_s = hop / fs;
    const double ts = ts_in[c];
    const double* tsp = tsc + f0 + c;          // timestamp of every position 0..frames (dsd_times_kernel)
    double ts_cur = ts;
    long long pos = 0;
    int best_cnt = 0, best_idx = 0;
    bool raining = true;
    auto clear_all = [&]() {
        for (int i = 0; i < DSD_OUT; i++) energy[i] = 0.0;
        for (int i = 0; i < p.n_bins; i++) { peak_hist[i] = 0; freq_hist[i] = 0.0; }
        best_cnt = 0; best_idx = 0;
    };
    auto remaining = [&]() -> long long { return (long long)n - pos * p.hop; };
    auto tti = [&]() -> double {
        double t = 60.0 - fmod(ts_cur, 60.0);
        if (t < hop_s) t += 60.0;
        return t;
    };
    auto do_frame = [&]() {
        const int64_t i = f0 + pos;
        const double pe = pk_val[i];
        const int pi = pk_idx[i];
        if (pe != 0.0) {
            const int cnt = ++peak_hist[pi];
            freq_hist[pi] += pe;
            if (cnt > best_cnt || (cnt == best_cnt && pi < best_idx)) { best_cnt = cnt; best_idx = pi; }
        }
        const int a = aux[i];
        energy[32 + (a & 0xff)] = (double)best_idx;
        if (a & 0x100) { for (int k = 0; k < p.n_bins; k++) peak_hist[k] = 0; best_cnt = 0; best_idx = 0; }
        const int h1 = (a >> 16) & 0xff;
        if (h1) energy[h1 - 1] += 1.0;
        pos++;
        ts_cur = tsp[pos];
    };
    const int num_minutes = (int)ceil((double)n / (fs * 60.0));
    for (int m = 0; m < num_minutes && produced < p.max_minutes; m++) {
        clear_all();
        bool alive = true;
        if (raining) {
            long long frames = (long long)(tti() * fs / hop);
            const long long fr_rem = (long long)((double)remaining() / hop);
            if (fr_rem < frames) frames = fr_rem;
            if (remaining() < p.L) frames = 0;
            for (long long f = 0; f < frames; f++) if (remaining() >= p.L) do_frame();
            for (int i = 0; i < p.n_bins; i++) {
                int j = (int)(log(freq_hist[i] + 2.719) * 25.0);
                if (j > 255) j = 255;
                if (i >= p.lwin0 && i <= p.lwin1) energy[62 + i - p.lwin0] = (double)j;
                if (p.hwin0 != p.lwin1 && i >= p.hwin0 && i <= p.hwin1) energy[62 + (i - p.hwin0) + 19] = (double)j;
            }
        } else {
            const double check = ts_cur + tti() - 3.0;
            while (ts_cur < check) {
                pos++;
                ts_cur = tsp[pos];             // pos <= frames here: the loop leaves as soon as no whole frame remains
                if (remaining() < p.L) { alive = false; break; }
            }
            if (!alive) break;
            clear_all();
            while (ts_cur < check + 3.0) {
                if (remaining() >= p.L) do_frame();
                else { alive = false; break; }
            }
            if (!alive) break;
        }
        raining = false;
        for (int i = 0; i < 32; i++) if (energy[i] != 0.0) raining = true;
        for (int i = 0; i < DSD_OUT; i++) o[(size_t)produced * DSD_OUT + i] = energy[i];
        produced++;
        if (remaining() < p.L) break;
    }
    n_minutes[c] = produced;
}

// The same machine, one WARP per clip.  The thread-per-clip kernel above waits for four uncoalesced global loads per
// frame and keeps its histograms in local memory (8.7 ms for 256 x 300 s clips on 256 threads).  Here the warp reads
// the per-frame quantities 32 positions at a time (lane l holds position window + l, one coalesced load each) and hands
// them round by shuffles; every lane carries the scalar state redundantly (uniform control flow); the histograms and
// the minute's 100-vector live in shared memory, where all lanes write identical values (or, for the clears and the
// per-bin loops, a lane-strided share followed by a warp barrier).  Same expressions, same order: bit-equal outputs.
__global__ void __launch_bounds__(32) dsd_minutes_warp_kernel(const __grid_constant__ DsdDev p, int n_clips, const int64_t* __restrict__ samp_off,
                                                              const int64_t* __restrict__ fr_off, const double* __restrict__ ts_in,
                                                              const double* __restrict__ tsc, const int* __restrict__ aux,
                                                              const int* __restrict__ pk_idx, const double* __restrict__ pk_val,
                                                              double* __restrict__ out, int* __restrict__ n_minutes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned FULL = 0xffffffffu;
    const int c = blockIdx.x, lane = threadIdx.x;
    if (c >= n_clips) return;
    const int64_t n = samp_off[c + 1] - samp_off[c];
    const int64_t f0 = fr_off[c];
    const long long nfr = fr_off[c + 1] - f0;
    double* o = out + (size_t)c * p.max_minutes * DSD_OUT;
    int produced = 0;
    if (n < p.L) { if (lane == 0) n_minutes[c] = 0; return; }
    double* freq_hist = reinterpret_cast<double*>(smem_raw);            // [n_bins]
    double* energy = freq_hist + p.n_bins;                              // [DSD_OUT]
    int* peak_hist = reinterpret_cast<int*>(energy + DSD_OUT);          // [n_bins]
    const double fs = (double)p.fs, hop = (double)p.hop;
    const double hop_s = hop / fs;
    const double ts = ts_in[c];
    const double* tsp = tsc + f0 + c;          // timestamp of every position 0..frames (dsd_times_kernel)
    // window of 32 positions: lane l holds the quantities of position wb + l
    long long wb = -1000;
    double w_ts = 0.0, w_pe = 0.0;
    int w_pi = 0, w_a = 0;
    auto ensure = [&](long long pos) {
        if (pos >= wb && pos < wb + 32) return;
        wb = pos;
        const long long q = pos + lane;
        w_ts = q <= nfr ? tsp[q] : 0.0;
        const bool in = q < nfr;
        w_pe = in ? pk_val[f0 + q] : 0.0;
        w_pi = in ? pk_idx[f0 + q] : 0;
        w_a = in ? aux[f0 + q] : 0;
    };
    auto ts_at = [&](long long pos) { ensure(pos); return __shfl_sync(FULL, w_ts, (int)(pos - wb)); };
    double ts_cur = ts;
    long long pos = 0;
    int best_cnt = 0, best_idx = 0;
    bool raining = true;
    auto clear_all = [&]() {
        __syncwarp();
        for (int i = lane; i < DSD_OUT; i += 32) energy[i] = 0.0;
        for (int i = lane; i < p.n_bins; i += 32) { peak_hist[i] = 0; freq_hist[i] = 0.0; }
        __syncwarp();
        best_cnt = 0; best_idx = 0;
    };
    auto remaining = [&]() -> long long { return (long long)n - pos * p.hop; };
    auto tti = [&]() -> double {
        double t = 60.0 - fmod(ts_cur, 60.0);
        if (t < hop_s) t += 60.0;
        return t;
    };
    auto do_frame = [&]() {
        ensure(pos);
        const int src = (int)(pos - wb);
        const double pe = __shfl_sync(FULL, w_pe, src);
        const int pi = __shfl_sync(FULL, w_pi, src);
        const int a = __shfl_sync(FULL, w_a, src);
        if (pe != 0.0) {
            const int cnt = peak_hist[pi] + 1;          // every lane writes the same values
            peak_hist[pi] = cnt;
            freq_hist[pi] += pe;
            if (cnt > best_cnt || (cnt == best_cnt && pi < best_idx)) { best_cnt = cnt; best_idx = pi; }
        }
        energy[32 + (a & 0xff)] = (double)best_idx;
        if (a & 0x100) {
            __syncwarp();
            for (int k = lane; k < p.n_bins; k += 32) peak_hist[k] = 0;
            __syncwarp();
            best_cnt = 0; best_idx = 0;
        }
        const int h1 = (a >> 16) & 0xff;
        if (h1) energy[h1 - 1] += 1.0;
        pos++;
        ts_cur = ts_at(pos);
    };
    const int num_minutes = (int)ceil((double)n / (fs * 60.0));
    for (int m = 0; m < num_minutes && produced < p.max_minutes; m++) {
        clear_all();
        bool alive = true;
        if (raining) {
            long long frames = (long long)(tti() * fs / hop);
            const long long fr_rem = (long long)((double)remaining() / hop);
            if (fr_rem < frames) frames = fr_rem;
            if (remaining() < p.L) frames = 0;
            for (long long f = 0; f < frames; f++) if (remaining() >= p.L) do_frame();
            __syncwarp();
            for (int i = lane; i < p.n_bins; i += 32) {
                int j = (int)(log(freq_hist[i] + 2.719) * 25.0);
                if (j > 255) j = 255;
                if (i >= p.lwin0 && i <= p.lwin1) energy[62 + i - p.lwin0] = (double)j;
                if (p.hwin0 != p.lwin1 && i >= p.hwin0 && i <= p.hwin1) energy[62 + (i - p.hwin0) + 19] = (double)j;
            }
            __syncwarp();
        } else {
            const double check = ts_cur + tti() - 3.0;
            while (ts_cur < check) {
                pos++;
                ts_cur = ts_at(pos);           // pos <= frames here: the loop leaves as soon as no whole frame remains
                if (remaining() < p.L) { alive = false; break; }
            }
            if (!alive) break;
            clear_all();
            while (ts_cur < check + 3.0) {
                if (remaining() >= p.L) do_frame();
                else { alive = false; break; }
            }
            if (!alive) break;
        }
        __syncwarp();
        raining = __ballot_sync(FULL, energy[lane] != 0.0) != 0u;       // any of the 32 drop-size bins
        for (int i = lane; i < DSD_OUT; i += 32) o[(size_t)produced * DSD_OUT + i] = energy[i];
        produced++;
        if (remaining() < p.L) break;
    }
    if (lane == 0) n_minutes[c] = produced;
}

}  // namespace apt
