// apt_dsd.cuh -- drop-size-distribution emulator: the compute path of the reference's
// transform.process_audio_file_dsd (transform.py:251-313), i.e. DsdProcessingEmualtor
// (host_analysis/device_dsd_processing_emulator.py:16-314), batched over clips.
//
//   dsd_frame_kernel    one CTA per hop position: (optionally Hann-windowed) frame -> float64 FFT ->
//                       |X[k]| -> drop energy (sum over the rain band), peak bin and its magnitude
//                       (process_audio_frame, :128-180, the spectral part)
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

// One thread per clip.  Every floating-point expression is evaluated in float64 in the reference's order
// (Python floats), because the comparisons against minute / rain-check boundaries depend on the rounding.
__global__ void dsd_minutes_kernel(const __grid_constant__ DsdDev p, int n_clips, const int64_t* __restrict__ samp_off,
                                   const int64_t* __restrict__ fr_off, const double* __restrict__ ts_in,
                                   const double* __restrict__ drop, const int* __restrict__ pk_idx, const double* __restrict__ pk_val,
                                   double* __restrict__ out, int* __restrict__ n_minutes) {
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
    const double ts_start = ts - fmod(ts, 60.0);
    double ts_cur = ts;
    long long frame_count = (long long)(fmod(ts_cur, 60.0) * fs / hop);
    long long pos = 0;
    int best_cnt = 0, best_idx = 0;
    bool raining = true;
    const double logbase = log(p.rain_log_base);
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
        const int nxt = (int)(fmod(ts_cur + hop_s, 60.0) / 2.0);
        const int cur = (int)(fmod(ts_cur, 60.0) / 2.0);
        energy[32 + cur] = (double)best_idx;
        if (nxt != cur) { for (int k = 0; k < p.n_bins; k++) peak_hist[k] = 0; best_cnt = 0; best_idx = 0; }
        const double de = drop[i];
        if (de > p.rain_thr) {
            int h = (int)floor(log(1.0 + (de - p.rain_thr) * p.rain_log_factor) / logbase);
            h = h > 31 ? 31 : (h < 0 ? 0 : h);
            energy[h] += 1.0;
        }
        pos++; frame_count++;
        ts_cur = ts_start + (double)(frame_count * p.hop) / fs;
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
                pos++; frame_count++;
                ts_cur = ts_start + (double)(frame_count * p.hop) / fs;
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

}  // namespace apt
