// apt_kernels.cuh -- sm_100a kernels of the hot path.
//
// Kernel map (DESIGN.md has the data layout and rooflines):
//   stft256_kernel      K1+K2+K3  PCM -> frames -> Hann -> rFFT-256 -> power -> band planes
//   td_features_kernel  K6+K7     PCM -> zero-phase SOS prefilter -> crest / kurtosis / block features
//   trk1_kernel         K4        tracker pass 1 on the mode bins (serial in time)
//   flux_kernel         K5        dB normalisation, flux, per-mode sums (parallel)
//   base_kernel         K8        float64 quantile baselines + normalisation (serial in time)
//   decide/compact      K8+K9     decision, labels, confidences; ordered event indices
//   trk2_kernel         K4        tracker pass 2 gated by the labels (serial in time)
//   db_kernel           K9        noise-floor dB plane, sums, level-0 histogram (parallel)
//   select_*            K9        exact median of the noise-floor dB plane (3-level radix select)
//   finalize_kernel     K9        clip statistics rows
//
// Compile with -fmad=false: roundings are deliberate (see apt_math.cuh).
#pragma once
#include <cuda_runtime.h>
#include "apt_math.cuh"
#include "../../include/apt_b200.h"

namespace apt {

__constant__ uint32_t kSvmlLog10TabDev[64] = {
    0xbdc9ae9b, 0xbda6fcf4, 0xbd8bac76, 0xbd6bca30, 0xbd48a99b, 0xbd2c0a9f, 0xbd1480db, 0xbd00faf2,
    0xbe823aa9, 0xbe656348, 0xbe4afbb9, 0xbe346895, 0xbe20ffff, 0xbe103a0b, 0xbe01a91c, 0xbde9e84e,
    0x3e13d888, 0x3e10a87c, 0x3e0b95c3, 0x3e057f0b, 0x3dfde038, 0x3df080d9, 0x3de34c1e, 0x3dd68333,
    0x3dac6e8e, 0x3dd54a51, 0x3df30f40, 0x3e04235d, 0x3e0b7033, 0x3e102c90, 0x3e12ebad, 0x3e141ff8,
    0xbe5e5a9b, 0xbe5e2677, 0xbe5d83f5, 0xbe5c6016, 0xbe5abd0b, 0xbe58a6fd, 0xbe562e02, 0xbe5362f8,
    0xbe68e27c, 0xbe646747, 0xbe619a73, 0xbe5ff05a, 0xbe5f0570, 0xbe5e92d0, 0xbe5e662b, 0xbe5e5c08,
    0x3ede5bd8, 0x3ede5b45, 0x3ede57d8, 0x3ede4eb1, 0x3ede3d37, 0x3ede2166, 0x3eddf9d9, 0x3eddc5bb,
    0x3ede08ed, 0x3ede32e7, 0x3ede4967, 0x3ede5490, 0x3ede597f, 0x3ede5b50, 0x3ede5bca, 0x3ede5bd9};

#ifdef APT_PROFILE_PHASES
// phase stamps of one interior tile (profiling builds only: profiles/micro/*_phases.cu)
__device__ long long g_stamp[64];
#define APT_STAMP(i) do { __syncthreads(); if (threadIdx.x == 0 && blockIdx.x == 77 && blockIdx.y == 7) o.dbg[i] = clock64(); } while (0)
#define APT_STAMP2(i) do { __syncthreads(); if (threadIdx.x == 0 && blockIdx.x == 77 && blockIdx.y == 7) g_stamp[i] = clock64(); } while (0)
#else
#define APT_STAMP(i) do { } while (0)
#define APT_STAMP2(i) do { } while (0)
#endif

struct DevParams {
    int n_fft, hop, F, band_lo, K, M;
    int mode_lo[APT_MAX_MODES], mode_hi[APT_MAX_MODES];
    int mode_blo[APT_MAX_MODES], mode_bhi[APT_MAX_MODES];
    double mode_w[APT_MAX_MODES];
    float trk_eta, trk_alpha, trk_1m_alpha, trk_floor, trk_q, trk_nq, trk_maxr;
    double ema_up, ema_down;
    int adaptive_q; double aq_base, aq_min, aq_alpha;   // adaptive quantile of pass 2
    int pre_smooth, median;                              // pre_smooth_frames / median_frames (<= 1: off)
    int snr_gate; float snr1; float snr_pow; uint32_t snr_mask[4];      // spectral SNR gating of the oversubtraction
    int td_g;                                            // stride of the TD block statistics (128, or 64 when the hop asks for it)
    int bypass_cls;                                      // bypass_classifier: every frame NOISE
    int warm_need;
    float eps32;
    int use_norm, ratio_db;
    double bl_q, bl_eta, bl_alpha, bl_floor;
    int norm_enable;
    float norm_min;
    float thr0, thr1, thr2, thr3;
    int min_support;
    float gate_thr;
    int has_ku;
    float ku;
    float noise_hi, mf_noise_max;
    int n_sos, padlen;
    double sos[APT_MAX_SOS][6];
    double zi[APT_MAX_SOS][2];
    double eps64;
    int blk_len, blk_hop, blk_pp, blk_smooth;
    int low_lo, low_hi, rain_lo, rain_hi;
    double rolloff;
    int suppressor_bypass;
    int min_frames;
    // suppressor gain
    int gain_mode, adaptive_gain, gain_freq_smooth, n_gain_taps, use_lagged;
    float oversub_noise, oversub_rain, gain_floor, gain_ceil;
    float gain_taps[APT_MAX_GAIN_TAPS];
    float alpha_noise, om_noise, alpha_base, om_base, gain_eps;
    // optional peak-structure features
    int peak_top_p, primary_top_m;
    double peak_prom_db, peak_min_above_floor, peak_ratio_min;
    float peak_valid_prom_min, peak_valid_prom_max;
};

// A launch covers clips [clip0, clip0 + n_clips) of the plan; the offset arrays are the plan's
// (absolute) prefix arrays.
struct Batch {
    int clip0, n_clips;
    const int64_t* samp_off;   // [plan clips + 1]
    const int64_t* frame_off;  // [plan clips + 1]
    int tile0;                 // first tile of every clip this launch covers (tiled kernels)
    int ta, tb;                // clip-local frame range [ta, tb) this launch covers (serial / per-frame kernels)
    int ntile;                 // tiles of every clip this launch covers, from tile0 (persistent tiled kernels: a CTA
                               // walks tile0 + blockIdx.x, + gridDim.x, ... below tile0 + ntile)
};

// tile -> clip: tiled kernels launch a 2-D grid, blockIdx.y = clip of the launch's range, blockIdx.x = tile
// inside the clip (grid.x = the largest tile count of the range; surplus CTAs of shorter clips exit at
// once).  tile_off is the plan's absolute prefix array of tiles per clip.  No search, no per-tile table.
// A launch may cover only the tiles [b.tile0, b.tile0 + gridDim.x) of every clip (one time segment of the pipelined run).
__device__ __forceinline__ bool tile_clip(const Batch& b, const int64_t* __restrict__ tile_off, int& c, int64_t& tile_in_clip) {
    c = b.clip0 + (int)blockIdx.y;
    tile_in_clip = (int64_t)b.tile0 + blockIdx.x;
    return tile_in_clip < __ldg(tile_off + c + 1) - __ldg(tile_off + c);
}

// ---------------------------------------------------------------------------------------------
// PCM access: int16 (wire format) or float32 (what the reference's loader hands to processors)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float load_sample(const int16_t* p, int64_t i) { return pcm_to_f32(__ldg(p + i)); }
__device__ __forceinline__ float load_sample(const float* p, int64_t i) { return __ldg(p + i); }

// Vector width of the staging loads in samples: 4 either way (8-byte loads of int16, 16-byte loads of float32),
// so that one load feeds exactly one 128-bit shared-memory store and the stores of a warp are contiguous.
// First i >= 0 such that element g + i of the PCM buffer sits on that vector's boundary:
template <typename PCM>
__device__ __forceinline__ int run_head(int64_t g) { return (int)((4 - (g % 4 + 4) % 4) % 4); }

// Loads the n samples that start at absolute element index g (all inside the PCM buffer) and hands
// them to put(i, value) / put4(i, four values i..i+3).  Aligned vector loads; every load of a thread is
// issued before the first conversion so the misses overlap (MAXV = vectors per thread upper bound).
// put4 is called with i = run_head(g) (mod 4) only, so a destination shifted by (4 - run_head % 4) % 4
// floats takes aligned 128-bit shared-memory stores.
// RAW: int16 samples are handed over unscaled ((float)s instead of s / 32767): for consumers that are invariant to the
// scale or undo it later (the float32 fast path of the TD gate).
template <bool RAW> __device__ __forceinline__ float cvt_pcm(int16_t s) { return RAW ? (float)s : pcm_to_f32(s); }
template <bool RAW> __device__ __forceinline__ float load_in(const int16_t* p, int64_t i) { return cvt_pcm<RAW>(__ldg(p + i)); }
template <bool RAW> __device__ __forceinline__ float load_in(const float* p, int64_t i) { return __ldg(p + i); }
// NT: the CTA size when the caller knows it at compile time (the vector loads of a thread then sit at immediate offsets
// from one pointer instead of a 64-bit multiply-add each), 0: blockDim.x.
template <int MAXV, typename PCM, typename Put, typename Put4, bool RAW = false, int NT = 0>
__device__ __forceinline__ void load_run(const PCM* __restrict__ pcm, int64_t g, int n, Put put, Put4 put4) {
    constexpr int VEC = 4;
    const int tid = threadIdx.x, nt = NT ? NT : (int)blockDim.x;
    int head = run_head<PCM>(g);
    if (head > n) head = n;
    const int nvec = (n - head) / VEC;
    const int tail0 = head + nvec * VEC;
    if constexpr (sizeof(PCM) == 2) {
        int2 raw[MAXV];
        const int2* vp = reinterpret_cast<const int2*>(pcm + g + head) + tid;
#pragma unroll
        for (int r = 0; r < MAXV; r++) {
            const int v = tid + r * nt;
            if (v < nvec) raw[r] = __ldg(vp + r * nt);
        }
        if (tid < head) put(tid, load_in<RAW>(pcm, g + tid));
        if (tail0 + tid < n) put(tail0 + tid, load_in<RAW>(pcm, g + tail0 + tid));
#pragma unroll
        for (int r = 0; r < MAXV; r++) {
            const int v = tid + r * nt;
            if (v < nvec) {
                const int2 q = raw[r];
                put4(head + v * VEC, make_float4(cvt_pcm<RAW>((int16_t)(q.x & 0xffff)), cvt_pcm<RAW>((int16_t)(q.x >> 16)),
                                                 cvt_pcm<RAW>((int16_t)(q.y & 0xffff)), cvt_pcm<RAW>((int16_t)(q.y >> 16))));
            }
        }
    } else {
        float4 raw[MAXV];
        const float4* vp = reinterpret_cast<const float4*>(pcm + g + head) + tid;
#pragma unroll
        for (int r = 0; r < MAXV; r++) {
            const int v = tid + r * nt;
            if (v < nvec) raw[r] = __ldg(vp + r * nt);
        }
        if (tid < head) put(tid, load_sample(pcm, g + tid));
        if (tail0 + tid < n) put(tail0 + tid, load_sample(pcm, g + tail0 + tid));
#pragma unroll
        for (int r = 0; r < MAXV; r++) {
            const int v = tid + r * nt;
            if (v < nvec) put4(head + v * VEC, raw[r]);
        }
    }
    // vectors beyond MAXV per thread (never for the tile sizes in this file)
    for (int v = tid + MAXV * nt; v < nvec; v += nt) {
        const int i = head + v * VEC;
        for (int e = 0; e < VEC; e++) put(i + e, load_in<RAW>(pcm, g + i + e));
    }
}

// Stages samples [s0, s0+n) of a clip (clip-relative indices, zero outside [0, N)) into shared memory
// as float32.  The body uses 128-bit global loads on 16-byte aligned addresses.
struct IdentityPos { __device__ __forceinline__ int operator()(int i) const { return i; } };
template <typename PCM, typename Pos = IdentityPos>
__device__ __forceinline__ void stage_clip_f32(const PCM* __restrict__ pcm, int64_t clip_base, int64_t N,
                                               int64_t s0, int n, float* __restrict__ dst, Pos pos = Pos()) {
    constexpr int VEC = 16 / (int)sizeof(PCM);
    const int tid = threadIdx.x, nt = blockDim.x;
    // global element index of dst[0]
    const int64_t g0 = clip_base + s0;
    // first i such that (g0 + i) is a multiple of VEC
    int head = (int)((VEC - (g0 % VEC + VEC) % VEC) % VEC);
    if (head > n) head = n;
    for (int i = tid; i < head; i += nt) {
        int64_t s = s0 + i;
        dst[pos(i)] = (s >= 0 && s < N) ? load_sample(pcm, clip_base + s) : 0.0f;
    }
    const int nvec = (n - head) / VEC;
    for (int v = tid; v < nvec; v += nt) {
        const int i = head + v * VEC;
        const int64_t s = s0 + i;
        if (s >= 0 && s + VEC <= N) {
            const int4 raw = __ldg(reinterpret_cast<const int4*>(pcm + clip_base + s));
            if constexpr (sizeof(PCM) == 2) {
                const int16_t* h = reinterpret_cast<const int16_t*>(&raw);
#pragma unroll
                for (int e = 0; e < 8; e++) dst[pos(i + e)] = pcm_to_f32(h[e]);
            } else {
                const float* f = reinterpret_cast<const float*>(&raw);
#pragma unroll
                for (int e = 0; e < 4; e++) dst[pos(i + e)] = f[e];
            }
        } else {
#pragma unroll
            for (int e = 0; e < VEC; e++) {
                int64_t se = s + e;
                dst[pos(i + e)] = (se >= 0 && se < N) ? load_sample(pcm, clip_base + se) : 0.0f;
            }
        }
    }
    for (int i = head + nvec * VEC + tid; i < n; i += nt) {
        int64_t s = s0 + i;
        dst[pos(i)] = (s >= 0 && s < N) ? load_sample(pcm, clip_base + s) : 0.0f;
    }
}

// ---------------------------------------------------------------------------------------------
// K1+K2+K3: STFT (n_fft = 256) + power + band planes
// ---------------------------------------------------------------------------------------------
constexpr int STFT_TF = 32;          // frames per tile
constexpr int STFT_NT = STFT_TF * 8; // 8 lanes per frame
constexpr int STFT_PS = 132;         // row stride of the power tile (F = 129)
constexpr int STFT_XS = (STFT_TF + 1) * 136 + 8;   // staged samples: 8 floats of padding per 128, + alignment shift
// Shared-memory area of one frame, in complex elements of the working precision.  It is used three times:
// (1) all areas together first hold the staged samples of the tile (dead once pass A has them in registers),
// (2) the frame's 128-element exchange buffer between the FFT passes, (3) the frame's complex64 spectrum
// (129 x 8 bytes from byte (t & 1) * 64: the two frames of a half-warp store their 64-byte runs into one
// wavefront) and, from byte STFT_POFF, its float32 power row.  Strides are multiples of 128 bytes.
template <typename T> struct ExFrame { static constexpr int value = sizeof(T) == 8 ? 128 : 224; };
constexpr int STFT_POFF = 1152;      // byte offset of the power row inside a frame area (after 64 + 129 * 8)

template <typename T>
struct FftTables {
    const T* win;          // [256]
    const cx<T>* tw128;    // [128]
    const cx<T>* tw256;    // [129]
};

struct StftOut {
    float* S;            // [nF][F][2]
    float* P;            // [nF][F]
    float* P_band;       // [nF][K]   (scratch plane consumed by clip_seq_kernel)
    float* band_energy;  // [M+1][nF]
    float* raw;          // [21][nF]
    const float* freqs;  // [F] device
    int64_t nF;
};

template <typename T>
constexpr size_t stft_smem_bytes() {
    static_assert(sizeof(float) * STFT_XS <= sizeof(cx<T>) * (size_t)STFT_TF * ExFrame<T>::value, "samples must fit the frame areas");
    static_assert(STFT_POFF + sizeof(float) * STFT_PS <= sizeof(cx<T>) * ExFrame<T>::value, "power row must fit the frame area");
    return sizeof(cx<T>) * (size_t)STFT_TF * ExFrame<T>::value + sizeof(T) * 256 + sizeof(cx<T>) * (128 + 130) + 64;
}

__device__ void raw_features_frame(const DevParams& p, const float* __restrict__ Pt, const float* __restrict__ freqs,
                                   float* __restrict__ out, int64_t stride);

template <typename T, typename PCM>
__global__ void __launch_bounds__(STFT_NT, 3) stft256_kernel(const __grid_constant__ DevParams p, Batch b,
                                                          const PCM* __restrict__ pcm,
                                                          const int64_t* __restrict__ tile_off, FftTables<T> tab,
                                                          StftOut o) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<T>* s_ex = reinterpret_cast<cx<T>*>(smem_raw);
    constexpr int EXF = ExFrame<T>::value;
    cx<T>* s_tw128 = s_ex + (size_t)STFT_TF * EXF;   // pass-A twiddles, [k1][lane]: W128^(lane * k1)
    cx<T>* s_tw256 = s_tw128 + 128;
    T* s_win = reinterpret_cast<T*>(s_tw256 + 130);
    float* s_x = reinterpret_cast<float*>(smem_raw);   // staged samples overlay the frame areas until pass A holds them
    auto S_row = [&](int t) { return reinterpret_cast<float2*>(s_ex + (size_t)t * EXF) + (t & 1) * 8; };
    auto P_row = [&](int t) { return reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(s_ex + (size_t)t * EXF) + STFT_POFF); };

    const int tid = threadIdx.x;
    const int c = b.clip0 + (int)blockIdx.y;
    const int n_tiles = (int)(__ldg(tile_off + c + 1) - __ldg(tile_off + c));
    const int tile_end = (int)min((int64_t)n_tiles, (int64_t)b.tile0 + (int64_t)b.ntile);
    if (b.tile0 + (int)blockIdx.x >= tile_end) return;
    const int64_t base = __ldg(b.samp_off + c);
    const int64_t N = __ldg(b.samp_off + c + 1) - base;
    const int64_t f0 = __ldg(b.frame_off + c);
    const int T_clip = (int)(__ldg(b.frame_off + c + 1) - f0);

    APT_STAMP2(20);
    // tables once per CTA; the CTA then walks its tiles of the clip
    for (int i = tid; i < 256; i += STFT_NT) s_win[i] = tab.win[i];
    for (int i = tid; i < 128; i += STFT_NT) s_tw128[i] = tab.tw128[((i & 7) * (i >> 3)) & 127];
    for (int i = tid; i < 129; i += STFT_NT) s_tw256[i] = tab.tw256[i];
    for (int tile = b.tile0 + (int)blockIdx.x; tile < tile_end; tile += (int)gridDim.x) {
    const int t0 = tile * STFT_TF;
    const int nfr = min(STFT_TF, T_clip - t0);
    // Sample u of the tile (u = 0 is 128 samples before the first frame) lives at u + sh + padk * (u / 128):
    // at hop 128 the four frames of a warp then read different banks (padk = 8); sh aligns the 128-bit
    // staging stores.
    const int padk = p.hop == 128 ? 8 : 0;
    int sh = 0;
    {
        // frames t0 .. t0+TF-1 of hop p.hop (<= 128) cover samples [t0*hop - 128, (t0+TF-1)*hop + 128)
        const int64_t s0 = (int64_t)t0 * p.hop - 128;
        constexpr int NS_MAX = (STFT_TF + 1) * 128;
        const int NS_ = (STFT_TF - 1) * p.hop + 256;
        if (s0 >= 0 && s0 + NS_ <= N) {
            sh = (4 - (run_head<PCM>(base + s0) & 3)) & 3;
            auto xpos = [&](int u) { return u + sh + padk * (u >> 7); };
            auto put1 = [&](int i, float v) { s_x[xpos(i)] = v; };
            auto put4 = [&](int i, float4 v) {
                if ((i & 127) > 124) { s_x[xpos(i)] = v.x; s_x[xpos(i + 1)] = v.y; s_x[xpos(i + 2)] = v.z; s_x[xpos(i + 3)] = v.w; }
                else *reinterpret_cast<float4*>(s_x + xpos(i)) = v;
            };
            load_run<(NS_MAX / 4 + STFT_NT - 1) / STFT_NT + 1, PCM, decltype(put1), decltype(put4), false, STFT_NT>(pcm, base + s0, NS_, put1, put4);
        } else {
            stage_clip_f32(pcm, base, N, s0, NS_, s_x, [&](int u) { return u + padk * (u >> 7); });
        }
    }
    __syncthreads();
    APT_STAMP2(21);

    const int fr = tid >> 3, lane = tid & 7;
    const float* xs = s_x + sh + fr * (p.hop + padk);   // sample n of the frame: xs[n + (n >= 128 ? padk : 0)]
    cx<T>* ex = s_ex + (size_t)fr * EXF;
    // all STFT_TF frame slots run both passes (slots past the clip end transform zero padding and are never
    // written out): the warp barrier inside pass B needs every lane
    rfft256_passA<T>(lane, [&](int n) { return xs[n + (n >= 128 ? padk : 0)]; }, s_win, s_tw128, ex, [&]() { __syncthreads(); });
    __syncthreads();
    APT_STAMP2(22);
    // pass B: every lane emits its bins as complex64 into the frame's own (now consumed) exchange area
    {
        float2* sS = S_row(fr);
        // the default path reads the operating band only: bins no lane needs are neither unpacked nor stored
        const bool all_bins = o.S || o.P || o.raw || o.band_energy;
        rfft256_passB<T>(lane, ex, s_tw256,
                         [&](int k, T re, T im) { sS[k] = make_float2(d2f((double)re), d2f((double)im)); },
                         [&]() { __syncwarp(); }, all_bins ? 0 : p.band_lo, all_bins ? 128 : p.band_lo + p.K - 1);
    }
    APT_STAMP2(23);

    // power |S|^2 with numpy's complex64 abs
    const int64_t fbase = f0 + t0;
    const bool need_full = o.P || o.raw || o.band_energy;
    if (!need_full) {
        // default path (band plane only): the frame's 8 lanes take its band bins lane, lane + 8, ... -- fixed mapping, no
        // index arithmetic; the bins were written by other lanes of the same frame (same warp) in pass B
        __syncwarp();
        if (fr < nfr && o.P_band) {
            const float2* sS = S_row(fr) + p.band_lo;
            float* dst = o.P_band + (fbase + fr) * p.K;
            const int K = p.K;
#pragma unroll 3
            for (int kb = lane; kb < K; kb += 8) {
                const float2 z = sS[kb];
                const float a = np_cabsf_fast(z.x, z.y);
                dst[kb] = a * a;
            }
        }
        if (o.S) {
            __syncthreads();
            float2* dstS = reinterpret_cast<float2*>(o.S) + fbase * p.F;
            for (int i = tid; i < nfr * p.F; i += STFT_NT) {
                const int t = i / p.F, k = i - t * p.F;
                dstS[i] = S_row(t)[k];
            }
        }
        __syncthreads();      // the frame areas are restaged by the next tile
        continue;
    }
    __syncthreads();
    {
        const int klo = 0, nk = p.F;
        float* dstb = o.P_band ? o.P_band + fbase * p.K : nullptr;
        // all (frame, bin) pairs of the tile, flattened over the CTA: element e -> frame e / nk by a multiply with
        // ceil(2^24 / nk) (exact for e < 2^12, nk <= 129)
        const unsigned inv = (1u << 24) / (unsigned)nk + 1u;
        const int n_el = nfr * nk;
#pragma unroll 4
        for (int e = tid; e < n_el; e += STFT_NT) {
            const int t = (int)(((unsigned long long)(unsigned)e * inv) >> 24);
            const int k = klo + (e - t * nk);
            const float2 z = S_row(t)[k];
            const float a = np_cabsf_fast(z.x, z.y);
            const float pw = a * a;
            P_row(t)[k] = pw;
            const int kb = k - p.band_lo;
            if (dstb && kb >= 0 && kb < p.K) dstb[t * p.K + kb] = pw;
        }
        if (o.S) {
            float2* dstS = reinterpret_cast<float2*>(o.S) + fbase * p.F;
            for (int i = tid; i < nfr * p.F; i += STFT_NT) {
                const int t = i / p.F, k = i - t * p.F;
                dstS[i] = S_row(t)[k];
            }
        }
    }
    __syncthreads();
    if (o.P) {
        float* dst = o.P + fbase * p.F;
        for (int i = tid; i < nfr * p.F; i += STFT_NT) {
            int t = i / p.F, k = i - t * p.F;
            dst[i] = P_row(t)[k];
        }
    }
    if (o.band_energy) {
        for (int i = tid; i < nfr * (p.M + 1); i += STFT_NT) {
            int m = i / nfr, t = i - m * nfr;
            const float* Pt = P_row(t);
            double s = 0.0;
            if (m < p.M) {
                for (int k = p.mode_lo[m]; k <= p.mode_hi[m]; k++) s += (double)Pt[k];
            } else {
                for (int k = 0; k < p.K; k++) s += (double)Pt[p.band_lo + k];
                s += p.eps64;
            }
            o.band_energy[(int64_t)m * o.nF + fbase + t] = d2f(s);
        }
    }
    if (o.raw) {
        if (tid < nfr) raw_features_frame(p, P_row(tid), o.freqs, o.raw + fbase + tid, o.nF);
    }
    __syncthreads();      // the frame areas are restaged by the next tile
    }   // tiles of this CTA
    APT_STAMP2(24);
}

// ---------------------------------------------------------------------------------------------
// K1+K2+K3 for any power-of-two frame size 256..4096 and any hop (BASELINE config 5, features stage):
// one CTA per frame.  The real frame is packed into n_fft/2 complex points, transformed by a
// shared-memory radix-2 Stockham FFT (autosort, ping-pong buffers, twiddles W_N^k from a table), and
// unpacked to the n_fft/2+1 rfft bins.  Arithmetic type T as in stft256_kernel (float64 = reference).
// ---------------------------------------------------------------------------------------------
constexpr int STFTG_NT = 256;
constexpr int STFTG_PTS = 2048;      // complex points a CTA transforms at once: STFTG_PTS / (n_fft / 2) frames
template <typename T>
struct FftTablesG {
    const T* win;        // [n_fft]
    const cx<T>* tw;     // [n_fft]  W_N^k for every k (the Stockham passes index it up to (R-1)/R of the circle)
};
inline int stftg_frames_per_cta(int n_fft) { return 2 * STFTG_PTS / n_fft; }   // 256 -> 16, 512 -> 8, ..., 4096 -> 1
__host__ __device__ __forceinline__ int stftg_pos(int p) { return p + (p >> 3); }   // one pad element per 8: the stride-8 stores of a pass spread over the banks
template <typename T>
inline size_t stftg_smem_bytes(int n_fft) {
    return sizeof(cx<T>) * (size_t)(STFTG_PTS + STFTG_PTS / 8) + sizeof(float) * (size_t)(STFTG_PTS + 4 * stftg_frames_per_cta(n_fft) + 64);
}

__device__ __forceinline__ cx<double> ldg_cx(const cx<double>* p) { const double2 v = __ldg(reinterpret_cast<const double2*>(p)); return {v.x, v.y}; }
__device__ __forceinline__ cx<float> ldg_cx(const cx<float>* p) { const float2 v = __ldg(reinterpret_cast<const float2*>(p)); return {v.x, v.y}; }
// 128-bit (double) / 64-bit (float) shared-memory accesses of one complex element (cx<T> itself carries no alignment)
__device__ __forceinline__ cx<double> ld_cx(const cx<double>* p) { const double2 v = *reinterpret_cast<const double2*>(p); return {v.x, v.y}; }
__device__ __forceinline__ cx<float> ld_cx(const cx<float>* p) { const float2 v = *reinterpret_cast<const float2*>(p); return {v.x, v.y}; }
__device__ __forceinline__ void st_cx(cx<double>* p, cx<double> v) { *reinterpret_cast<double2*>(p) = make_double2(v.x, v.y); }
__device__ __forceinline__ void st_cx(cx<float>* p, cx<float> v) { *reinterpret_cast<float2*>(p) = make_float2(v.x, v.y); }

// one radix-R Stockham butterfly of stride q on the registers a[r] = x[i + r * H / R]; results to x[j + r * q]
template <typename T, int R>
__device__ __forceinline__ void stftg_butterfly(cx<T>* a, int i, int q, int twstep, const cx<T>* __restrict__ tw, cx<T>* __restrict__ x) {
    const int k = i & (q - 1);
    const int j = (i - k) * R + k;
    if (q > 1) {
        const cx<T>* twk = tw + k * twstep;
#pragma unroll
        for (int r = 1; r < R; r++) a[r] = cmul(a[r], ldg_cx(twk + (r - 1) * k * twstep));   // W_{Rq}^{r k}
    }
    if constexpr (R == 8) fft8(a);
    else if constexpr (R == 4) fft4(a[0], a[1], a[2], a[3]);
    else { const cx<T> u = a[0]; a[0] = cadd(u, a[1]); a[1] = csub(u, a[1]); }
#pragma unroll
    for (int r = 0; r < R; r++) st_cx(x + stftg_pos(j + r * q), a[r]);
}
// one in-place pass: every thread first takes the eight points of its 8 / R butterflies into registers, the CTA meets at a
// barrier, then the results overwrite the frame
template <typename T, int R>
__device__ __forceinline__ void stftg_pass(cx<T>* __restrict__ x, int tif, int H, int q, int N, const cx<T>* __restrict__ tw) {
    constexpr int NB = 8 / R;
    const int m = H / R, tpf = H >> 3, twstep = N / (R * q);
    cx<T> a[NB][R];
#pragma unroll
    for (int nb = 0; nb < NB; nb++)
#pragma unroll
        for (int r = 0; r < R; r++) a[nb][r] = ld_cx(x + stftg_pos(tif + nb * tpf + r * m));
    __syncthreads();
#pragma unroll
    for (int nb = 0; nb < NB; nb++) stftg_butterfly<T, R>(a[nb], tif + nb * tpf, q, twstep, tw, x);
    __syncthreads();
}

// STFT of any power-of-two frame size 256..4096 at any hop: the real FFT of a frame is a complex FFT of H = n_fft / 2
// points (even samples real, odd samples imaginary) plus the unpack.  H / 8 threads own a frame, a CTA 2048 / H frames.
// Mixed-radix Stockham autosort passes (8, 8, ... then 4s), each butterfly in registers: the first pass reads its eight
// inputs (i + r * H / 8: consecutive threads -> consecutive sample pairs) straight from global memory, windowed; the
// later passes work in place on one padded shared-memory buffer.  Threads of frames beyond the clip end run the same
// code on their (unused) frame area and skip the global accesses, so every barrier is met without divergence.
template <typename T, typename PCM>
__global__ void __launch_bounds__(STFTG_NT, 4) stft_generic_kernel(const __grid_constant__ DevParams p, Batch b,
                                                                const PCM* __restrict__ pcm, FftTablesG<T> tab, StftOut o, int fpc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = p.n_fft, H = N >> 1, F = H + 1, PS = H + 4;
    constexpr int BUF = STFTG_PTS + STFTG_PTS / 8;
    cx<T>* buf = reinterpret_cast<cx<T>*>(smem_raw);
    float* s_P = reinterpret_cast<float*>(buf + BUF);   // [fpc][PS]
    const int tid = threadIdx.x;
    const int c = b.clip0 + (int)blockIdx.y;
    const int64_t base = __ldg(b.samp_off + c);
    const int64_t Ns = __ldg(b.samp_off + c + 1) - base;
    const int64_t f0 = __ldg(b.frame_off + c);
    const int T_clip = (int)(__ldg(b.frame_off + c + 1) - f0);
    const int t0 = (b.tile0 + (int)blockIdx.x) * fpc;
    if (t0 >= T_clip) return;
    const int nfr = min(fpc, T_clip - t0);
    const int lgt = 31 - __clz(H >> 3);           // threads per frame = H / 8 = 1 << lgt
    const int tpf = 1 << lgt;
    const int f = tid >> lgt, tif = tid & (tpf - 1);
    const bool live = f < nfr;
    cx<T>* x = buf + f * (H + (H >> 3));          // the frame's area: H points + their pads
    // ---- pass 1 (radix 8, q = 1: no twiddles), inputs from global memory; centre padding with zeros (librosa center=True)
    {
        cx<T> a[8];
        const int64_t start = (int64_t)(t0 + f) * p.hop - H;       // clip-relative index of the frame's first sample
        const bool inside = start >= 0 && start + N <= Ns;
        const PCM* src = pcm + base + start;
        const cx<T>* win2 = reinterpret_cast<const cx<T>*>(tab.win);
        bool pair_ok = false;
        if constexpr (sizeof(PCM) == 2) pair_ok = live && inside && ((reinterpret_cast<uintptr_t>(src) & 3) == 0);
        float xe[8], xo[8];
        if (pair_ok) {          // the common case: whole frame inside the clip, sample pairs 32-bit aligned
            if constexpr (sizeof(PCM) == 2) {
                const uint32_t* src2 = reinterpret_cast<const uint32_t*>(src) + tif;
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    const uint32_t w = __ldg(src2 + r * tpf);
                    xe[r] = pcm_to_f32((int16_t)(w & 0xffffu)); xo[r] = pcm_to_f32((int16_t)(w >> 16));
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int n = tif + r * tpf;
                const int64_t sa = start + 2 * n;
                xe[r] = (live && sa >= 0 && sa < Ns) ? load_sample(src, 2 * n) : 0.0f;
                xo[r] = (live && sa + 1 >= 0 && sa + 1 < Ns) ? load_sample(src, 2 * n + 1) : 0.0f;
            }
        }
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const cx<T> w = ldg_cx(win2 + tif + r * tpf);
            a[r] = {w.x * (T)xe[r], w.y * (T)xo[r]};
        }
        stftg_butterfly<T, 8>(a, tif, 1, 0, tab.tw, x);
    }
    __syncthreads();
    // ---- remaining passes; schedule 8,4,4 | 8,8,4 | 8,8,8 | 8,8,4,4 | 8,8,8,4 for H = 128 .. 2048
    for (int q = 8; q < H;) {
        const int rem = H / q;                       // product of the radices still to do
        if (rem % 8 == 0 && rem != 16) { stftg_pass<T, 8>(x, tif, H, q, N, tab.tw); q *= 8; }
        else if (rem % 4 == 0) { stftg_pass<T, 4>(x, tif, H, q, N, tab.tw); q *= 4; }
        else { stftg_pass<T, 2>(x, tif, H, q, N, tab.tw); q *= 2; }
    }
    // ---- real-FFT unpack: X[k] = E + W_N^k O with E = (Z[k] + conj(Z[H-k]))/2, O = (Z[k] - conj(Z[H-k]))/(2i).
    // Thread tif takes k = tif + r * H/8, r = 0..7 (k = 0 comes out of the same formula with Z[H] = Z[0]); k = H apart.
    const T half = (T)0.5;
    if (live) {
        const int64_t fr = f0 + t0 + f;
        float* Pt = s_P + f * PS;
        float2* Sg = o.S ? reinterpret_cast<float2*>(o.S) + fr * F : nullptr;
        float* Pg = o.P ? o.P + fr * F : nullptr;
        float* Pb = o.P_band ? o.P_band + fr * p.K - p.band_lo : nullptr;
        const int bhi = p.band_lo + p.K;
        const bool keep = o.band_energy || o.raw;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int k = tif + r * tpf;
            const cx<T> zk = ld_cx(x + stftg_pos(k)), cn = cconj(ld_cx(x + stftg_pos((H - k) & (H - 1))));
            const cx<T> e = {(zk.x + cn.x) * half, (zk.y + cn.y) * half};
            const cx<T> d = csub(zk, cn);
            const cx<T> od = {d.y * half, -d.x * half};
            const cx<T> wo = cmul(od, ldg_cx(tab.tw + k));
            const float sr = d2f((double)(e.x + wo.x)), si = k == 0 ? 0.0f : d2f((double)(e.y + wo.y));
            if (Sg) Sg[k] = make_float2(sr, si);
            const float av = np_cabsf_fast(sr, si);
            const float pw = av * av;
            if (keep) Pt[k] = pw;
            if (Pg) Pg[k] = pw;
            if (Pb && k >= p.band_lo && k < bhi) Pb[k] = pw;
        }
        if (tif == 0) {   // k = H (Nyquist): real
            const cx<T> z0 = ld_cx(x);
            const float sr = d2f((double)(z0.x - z0.y));
            if (Sg) Sg[H] = make_float2(sr, 0.0f);
            const float av = np_cabsf_fast(sr, 0.0f);
            const float pw = av * av;
            if (keep) Pt[H] = pw;
            if (Pg) Pg[H] = pw;
            if (Pb && H >= p.band_lo && H < bhi) Pb[H] = pw;
        }
    }
    if (!(o.band_energy || o.raw)) return;
    __syncthreads();
    if (o.band_energy) {
        // one warp per frame, row after row: float64 partial sums per lane, combined by shuffles (tolerance feature)
        const int warp = tid >> 5, lane = tid & 31;
        for (int ff = warp; ff < nfr; ff += STFTG_NT / 32) {
            const float* Pt = s_P + ff * PS;
            float* be = o.band_energy + f0 + t0 + ff;
#pragma unroll 1
            for (int m = 0; m <= p.M; m++) {
                const int lo = m < p.M ? p.mode_lo[m] : p.band_lo, hi = m < p.M ? p.mode_hi[m] : p.band_lo + p.K - 1;
                double sacc = 0.0;
                for (int k = lo + lane; k <= hi; k += 32) sacc += (double)Pt[k];
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, d);
                if (lane == 0) be[(int64_t)m * o.nF] = d2f(m < p.M ? sacc : sacc + p.eps64);
            }
        }
    }
    if (o.raw && tid < nfr) raw_features_frame(p, s_P + tid * PS, o.freqs, o.raw + f0 + t0 + tid, o.nF);
}

// raw spectral features of one frame in float64 (feature_extraction.py:610-747)
__device__ void raw_features_frame(const DevParams& p, const float* __restrict__ Pt, const float* __restrict__ freqs,
                                   float* __restrict__ out, int64_t stride) {
    const double eps = p.eps64;
    const int lo = p.band_lo, K = p.K, F = p.F;
    double total = 0.0;
    for (int k = 1; k < F; k++) total += (double)Pt[k];
    const double total_nodc = F > 1 ? total + eps : (double)Pt[0] + eps;
    double op = 0.0, cen = 0.0;
    for (int k = 0; k < K; k++) { double v = (double)Pt[lo + k]; op += v; cen += (double)__ldg(freqs + lo + k) * v; }
    op += eps;
    cen /= op;
    double bw = 0.0;
    for (int k = 0; k < K; k++) { double d = (double)__ldg(freqs + lo + k) - cen; bw += d * d * (double)Pt[lo + k]; }
    bw = sqrt(bw / op);
    double lowr = 0.0, rainr = 0.0;
    for (int k = p.low_lo; k <= p.low_hi; k++) lowr += (double)Pt[k];
    for (int k = p.rain_lo; k <= p.rain_hi; k++) rainr += (double)Pt[k];
    lowr /= total_nodc; rainr /= total_nodc;
    double mbp[APT_MAX_MODES], mtot = 0.0;
    for (int i = 0; i < p.M; i++) {
        double s = 0.0;
        for (int k = p.mode_lo[i]; k <= p.mode_hi[i]; k++) s += (double)Pt[k];
        mbp[i] = s; mtot += s;
    }
    mtot += eps;
    double ent = 0.0, mean = 0.0, mx = -1e300;
    for (int i = 0; i < p.M; i++) {
        double r = mbp[i] / mtot;
        mbp[i] = r;
        ent += r * log(r + eps);
        mean += r;
        mx = r > mx ? r : mx;
    }
    mean /= (double)p.M;
    double var = 0.0;
    for (int i = 0; i < p.M; i++) { double d = mbp[i] - mean; var += d * d; }
    const double sd = sqrt(var / (double)p.M);
    double mlog = 0.0, marith = 0.0;
    for (int k = 0; k < K; k++) { double v = (double)Pt[lo + k] + eps; mlog += log(v); marith += v; }
    const double flat = exp(mlog / (double)K) / (marith / (double)K + eps);
    double frac = p.rolloff < 0.0 ? 0.0 : (p.rolloff > 1.0 ? 1.0 : p.rolloff);
    const double thr = frac * op;
    double cs = 0.0;
    int ridx = 0, found = 0, dom = 0;
    for (int k = 0; k < K; k++) {
        cs += (double)Pt[lo + k];
        if (!found && cs >= thr) { ridx = k; found = 1; }
        if (Pt[lo + k] > Pt[lo + dom]) dom = k;
    }
    const int ncep = 2 * (K - 1);
    double cep[5] = {0, 0, 0, 0, 0};
    if (K >= 2) {
        const double l0 = log(fmax((double)Pt[lo], eps)), lN = log(fmax((double)Pt[lo + K - 1], eps));
        for (int j = 0; j < 5; j++) cep[j] = l0 + ((j & 1) ? -lN : lN);
        for (int k = 1; k < K - 1; k++) {
            const double lg = 2.0 * log(fmax((double)Pt[lo + k], eps));
            for (int j = 0; j < 5; j++) cep[j] += lg * cospi(2.0 * (double)j * (double)k / (double)ncep);
        }
        for (int j = 0; j < 5; j++) cep[j] = (j < ncep) ? cep[j] / (double)ncep : 0.0;
    }
    out[0 * stride] = d2f(cen); out[1 * stride] = d2f(bw); out[2 * stride] = d2f(lowr); out[3 * stride] = d2f(rainr);
    for (int i = 0; i < 5; i++) out[(4 + i) * stride] = i < p.M ? d2f(mbp[i]) : 0.0f;
    out[9 * stride] = d2f(-ent); out[10 * stride] = d2f(sd); out[11 * stride] = d2f(mx); out[12 * stride] = d2f(flat);
    out[13 * stride] = __ldg(freqs + lo + ridx); out[14 * stride] = __ldg(freqs + lo + dom); out[15 * stride] = d2f(op);
    for (int j = 0; j < 5; j++) out[(16 + j) * stride] = d2f(cep[j]);
}

// ---------------------------------------------------------------------------------------------
// K6+K7: zero-phase SOS prefilter (scipy.signal.sosfiltfilt, float64) + TD frame features
// ---------------------------------------------------------------------------------------------
// A CTA filters one tile of TD_FT frames plus warm-up on both sides.  Every thread keeps its TD_CHUNK
// consecutive samples in REGISTERS through both filter directions (block-parallel IIR: direct pass from
// a zero state, Kogge-Stone scan of the chunk-final states with the chunk transition matrix, then the
// homogeneous response to the incoming state is added); shared memory only carries the staged PCM, the
// scan states and the float32 result.  The float64 pipe (64 FMA/clk/SM) is the bound of this kernel.
constexpr int TD_NT = 384;
constexpr int TD_FT = 56;       // TD frames per tile (fills the 384 x 23 sample buffer with 2 x 512 warm-up and the block-feature halo)
constexpr int TD_WARM = 512;    // warm-up samples each side (pole radius 0.928 -> < 2^-53 after 490)
constexpr int TD_CHUNK = 23;    // samples per thread (odd: conflict-free shared-memory walks)
constexpr int TD_LB = TD_NT * TD_CHUNK;   // filter buffer capacity of a tile
constexpr int TD_XF = TD_LB + TD_LB / 16 + 160;   // float32 staging / result area (padded layout)
constexpr int TD_MAXDIM = 2 * APT_MAX_SOS;

struct TdTables {
    // block-parallel IIR tables, computed at plan time for chunk length TD_CHUNK (state dim = 2*n_sos):
    const double* Alin;  // [dim*dim][32]  A^e for e = 0..31 (component-major), A = transition over one chunk
    const double* H;     // [chunk][dim]   output response at step n to a unit initial state
    int chunk;           // == TD_CHUNK
    int lb_max;          // == TD_LB
    int halo;            // extra valid samples each side of the frames (block features); <= 128
    int env_cap;         // capacity of the block-envelope scratch (doubles)
    // copies for state dimension <= 4 (n_sos <= 2) that travel in the kernel-parameter constant bank, so the
    // scan and the response pass use them as FMA operands without any load on the dependent path
    double Hc[TD_CHUNK * 4];   // [chunk][4]
    double Adc[5 * 16];        // [k][4][4]  A^(2^k), k = 0..4
    float Hcf[TD_CHUNK * 4];   // the same tables rounded to float32 (float32 fast path)
    float Adcf[5 * 16];
};
template <typename R> __device__ __forceinline__ R td_Hc(const TdTables& tb, int i);
template <> __device__ __forceinline__ double td_Hc<double>(const TdTables& tb, int i) { return tb.Hc[i]; }
template <> __device__ __forceinline__ float td_Hc<float>(const TdTables& tb, int i) { return tb.Hcf[i]; }
template <typename R> __device__ __forceinline__ R td_Adc(const TdTables& tb, int i);
template <> __device__ __forceinline__ double td_Adc<double>(const TdTables& tb, int i) { return tb.Adc[i]; }
template <> __device__ __forceinline__ float td_Adc<float>(const TdTables& tb, int i) { return tb.Adcf[i]; }

struct TdOut {
    float* td;       // [5][nF]  (rows: crest, kurtosis, block crest, block width, block post/pre)
    float* x_td;     // [nS] optional
    int64_t nF;
    int want_kurt;   // compute kurtosis
    int want_block;  // compute block features
    // Gate mode (default flags: the decision only consumes `crest > td_gate_threshold`).  gate != NULL: the kernel
    // writes the gate byte of every frame instead of the feature rows.  The float32 instantiation also appends the
    // (clip, tile) of every tile holding a frame whose crest factor lies within `guard` (relative) of the threshold
    // to `list`; the float64 instantiation launched in list mode then recomputes exactly those tiles and overwrites
    // their gate bytes, so the gate plane equals the float64 one bit for bit.
    uint8_t* gate;   // [nF]
    float* crest_dbg;// [nF] optional: the crest factor as the gate-mode kernel computed it (diagnostic plane)
    float guard;
    int2* list;      // [list_cap] flagged (clip, tile) pairs of this launch
    int* list_count;
    int list_cap;
    // Block mode (frame sizes other than 256 / 128).  blk_sum != NULL: the kernel writes the float32 sum of squares
    // (numpy's 8-accumulator order) and the peak of every 128-sample block of the prefiltered waveform instead of
    // frame features; see td_block_base for the layout.
    float* blk_sum;
    float* blk_max;
#ifdef APT_PROFILE_PHASES
    long long* dbg;  // [16] clock64 stamps of one interior tile (profiling builds only)
#endif
};


// first entry of clip c (whose samples start at `base`) in the block-statistics arrays: floor(base / 128) + c leaves every
// clip its floor(N / 128) entries (floor(a + b) >= floor(a) + floor(b)); the arrays hold nS / 128 + n_clips + 1 entries
__host__ __device__ __forceinline__ int64_t td_block_base(int64_t base, int c, int g = 128) { return base / g + c; }

inline size_t td_smem_bytes(int ns, int env_cap, size_t real_bytes = sizeof(double)) {
    return sizeof(float) * TD_XF + real_bytes * ((size_t)2 * ns * (TD_NT / 32) + 32 * 4 * ns * ns + (size_t)TD_CHUNK * 2 * ns) + 8 +
           sizeof(double) * 2 * (size_t)env_cap;
}

// one biquad cascade step (DF2T) in the working precision R, FMAs allowed (not bit-compared; 1e-16 level in float64)
template <int NS, typename R>
__device__ __forceinline__ R sos_step(const R (&c)[NS][6], R (&z)[NS][2], R x) {
#pragma unroll
    for (int s = 0; s < NS; s++) {
        R y = t_fma(c[s][0], x, z[s][0]);
        z[s][0] = t_fma(-c[s][4], y, t_fma(c[s][1], x, z[s][1]));
        z[s][1] = t_fma(-c[s][5], y, c[s][2] * x);
        x = y;
    }
    return x;
}

__device__ __forceinline__ double shfl_up_d(double v, int d) {
    return __hiloint2double(__shfl_up_sync(0xffffffffu, __double2hiint(v), d), __shfl_up_sync(0xffffffffu, __double2loint(v), d));
}
__device__ __forceinline__ double shfl_down_d(double v, int d) {
    return __hiloint2double(__shfl_down_sync(0xffffffffu, __double2hiint(v), d), __shfl_down_sync(0xffffffffu, __double2loint(v), d));
}
__device__ __forceinline__ float to_f32(double v) { return d2f(v); }
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float shfl_up_d(float v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ float shfl_down_d(float v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }

// One filter direction over the register-resident chunks.  Chunks are ordered by thread index (forward)
// or reverse thread index (REV); `ia` is the last thread holding samples.  The first chunk of the order
// starts from `zi * xe` when has_init, else from rest; every later chunk is full except the last chunk
// of the forward order.  Steps: (1) direct pass from a zero state, straight-line for full chunks;
// (2) scan of the chunk-final states over the chunk order: Kogge-Stone with shuffles inside each warp
// (v <- v + A^d v[pos-d]), then ONE shared-memory exchange adds A^(pos) times the previous warp's final
// state (older warps contribute A^32 and beyond, below 1e-20 by the plan-time decay check);
// (3) the homogeneous response to the incoming state is added to the chunk.
// s_Alin: [DIM*DIM][32] powers A^e, e = 0..31, component-major.  s_vend: [DIM][TD_NT/32].
template <int NS, bool REV, typename R>
__device__ __forceinline__ void iir_pass(const DevParams& p, const TdTables& tb, const R* __restrict__ s_Alin,
                                         const R* __restrict__ s_H, R (&y)[TD_CHUNK], int n_mine,
                                         bool has_init, R xe, int ia, R* __restrict__ s_vend) {
    constexpr int DIM = 2 * NS;
    constexpr int NW = TD_NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int wa = ia >> 5;
    const bool active = tid <= ia;
    // position inside the warp in chunk order, and whether an earlier warp exists
    const int la = min(31, ia - 32 * w);                 // last active lane of this warp (< 0: none)
    const int pos = REV ? la - lane : lane;
    const bool first_chunk = REV ? (tid == ia) : (tid == 0);
    const bool has_prev_warp = REV ? (w < wa) : (w > 0);
    R coef[NS][6];
#pragma unroll
    for (int s = 0; s < NS; s++)
#pragma unroll
        for (int j = 0; j < 6; j++) coef[s][j] = (R)p.sos[s][j];
    R z[NS][2];
    {
        const R x0 = (has_init && first_chunk) ? xe : (R)0;
#pragma unroll
        for (int s = 0; s < NS; s++) { z[s][0] = (R)p.zi[s][0] * x0; z[s][1] = (R)p.zi[s][1] * x0; }
    }
    if (n_mine == TD_CHUNK) {
#pragma unroll
        for (int jj = 0; jj < TD_CHUNK; jj++) {
            const int j = REV ? TD_CHUNK - 1 - jj : jj;
            y[j] = sos_step<NS, R>(coef, z, y[j]);
        }
    } else if (n_mine > 0) {
#pragma unroll
        for (int jj = 0; jj < TD_CHUNK; jj++) {
            const int j = REV ? TD_CHUNK - 1 - jj : jj;
            if (j < n_mine) y[j] = sos_step<NS, R>(coef, z, y[j]);
        }
    }
    APT_STAMP2(REV ? 10 : 0);
    // (2) scan.  v = chunk-final state (zero for threads without samples)
    R v[DIM];
#pragma unroll
    for (int r = 0; r < DIM; r++) v[r] = active ? z[r >> 1][r & 1] : (R)0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const int d = 1 << k;
        R u[DIM];
#pragma unroll
        for (int q = 0; q < DIM; q++) u[q] = REV ? shfl_down_d(v[q], d) : shfl_up_d(v[q], d);
        if (active && pos >= d) {
#pragma unroll
            for (int r = 0; r < DIM; r++)
#pragma unroll
                for (int q = 0; q < DIM; q++) {
                    const R a = (NS <= 2) ? td_Adc<R>(tb, k * 16 + r * DIM + q) : s_Alin[(r * DIM + q) * 32 + d];
                    v[r] = t_fma(a, u[q], v[r]);
                }
        }
    }
    APT_STAMP2(REV ? 11 : 1);
    // final state of this warp's last chunk in order, for the next warp
    if (REV ? (lane == 0) : (lane == 31)) {
#pragma unroll
        for (int r = 0; r < DIM; r++) s_vend[r * NW + w] = v[r];
    }
    // scanned state of the previous chunk inside the warp
    R sin_[DIM];
#pragma unroll
    for (int r = 0; r < DIM; r++) {
        const R nb = REV ? shfl_down_d(v[r], 1) : shfl_up_d(v[r], 1);
        sin_[r] = (pos >= 1) ? nb : (R)0;
    }
    __syncthreads();
    if (active && has_prev_warp && pos >= 0) {
        R ve[DIM];
#pragma unroll
        for (int q = 0; q < DIM; q++) ve[q] = s_vend[q * NW + (REV ? w + 1 : w - 1)];
#pragma unroll
        for (int r = 0; r < DIM; r++) {
            R arow[DIM];   // the row's loads are issued together, ahead of its FMA chain
#pragma unroll
            for (int q = 0; q < DIM; q++) arow[q] = s_Alin[(r * DIM + q) * 32 + pos];
#pragma unroll
            for (int q = 0; q < DIM; q++) sin_[r] = t_fma(arow[q], ve[q], sin_[r]);
        }
    }
    APT_STAMP2(REV ? 12 : 2);
    // (3) add the homogeneous response to the incoming state (every chunk but the first of the order)
    if (active && !first_chunk) {
        if (n_mine == TD_CHUNK) {
#pragma unroll
            for (int j = 0; j < TD_CHUNK; j++) {
                const int m = REV ? TD_CHUNK - 1 - j : j;
                if constexpr (sizeof(R) == 4) {
                    // float32 fast path: the response is accumulated straight onto the sample (one operation fewer)
                    R acc = y[j];
#pragma unroll
                    for (int r = 0; r < DIM; r++) acc = t_fma((NS <= 2) ? td_Hc<R>(tb, m * DIM + r) : s_H[m * DIM + r], sin_[r], acc);
                    y[j] = acc;
                } else {
                    R acc = (R)0;
#pragma unroll
                    for (int r = 0; r < DIM; r++) {
                        const R h = (NS <= 2) ? td_Hc<R>(tb, m * DIM + r) : s_H[m * DIM + r];
                        acc = (r == 0) ? h * sin_[0] : t_fma(h, sin_[r], acc);
                    }
                    y[j] += acc;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < TD_CHUNK; j++) {
                const int m = REV ? TD_CHUNK - 1 - j : j;
                if (j < n_mine) {
                    const R* h = s_H + m * DIM;
                    R acc = h[0] * sin_[0];
#pragma unroll
                    for (int r = 1; r < DIM; r++) acc = t_fma(h[r], sin_[r], acc);
                    y[j] += acc;
                }
            }
        }
    }
    __syncthreads();   // s_vend is reused by the next pass
    APT_STAMP2(REV ? 13 : 3);
}

// numpy pairwise sum of n = 128 * 2^m contiguous float32 values spread over 8 lanes (lane j = accumulator j).
// `ld(i)` returns element i.  All 8 lanes of the group return the total (0 + pairwise).
template <typename Load>
__device__ __forceinline__ float group8_np_sum(Load ld, int n, int lane, unsigned gmask) {
    float part[32];  // per 128-block results (n <= 4096)
    const int nblk = n >> 7;
    for (int blk = 0; blk < nblk; blk++) {
        float r = ld(blk * 128 + lane);
#pragma unroll
        for (int i = 1; i < 16; i++) r += ld(blk * 128 + 8 * i + lane);
        // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))
        float o1 = __shfl_xor_sync(gmask, r, 1);
        float s1 = (lane & 1) ? o1 + r : r + o1;
        float o2 = __shfl_xor_sync(gmask, s1, 2);
        float s2 = (lane & 2) ? o2 + s1 : s1 + o2;
        float o4 = __shfl_xor_sync(gmask, s2, 4);
        part[blk] = (lane & 4) ? o4 + s2 : s2 + o4;
    }
    // binary tree over blocks (n2 = n/2 is a multiple of 8 for these sizes)
    for (int w = 1; w < nblk; w <<= 1)
        for (int i = 0; i + w < nblk; i += 2 * w) part[i] = part[i] + part[i + w];
    return 0.0f + part[0];
}

template <typename X>
__device__ double td_peak_width_half(X x, int n, int peak) {
    int i = peak, lb = peak, rb = peak;
    double lmin = x(peak), rmin = x(peak);
    while (0 <= i && x(i) <= x(peak)) { if (x(i) < lmin) { lmin = x(i); lb = i; } i--; }
    i = peak;
    while (i <= n - 1 && x(i) <= x(peak)) { if (x(i) < rmin) { rmin = x(i); rb = i; } i++; }
    const double prom = x(peak) - (lmin > rmin ? lmin : rmin);
    const double height = x(peak) - prom * 0.5;
    i = peak;
    while (lb < i && height < x(i)) i--;
    double lip = (double)i;
    if (x(i) < height) lip += (height - x(i)) / (x(i + 1) - x(i));
    i = peak;
    while (i < rb && height < x(i)) i++;
    double rip = (double)i;
    if (x(i) < height) rip -= (height - x(i)) / (x(i - 1) - x(i));
    return rip - lip;
}

// float32 result layout: sample u (counted from 128 samples before the tile's first frame) lives at
// u + 8 * (u / 128): consecutive 128-sample blocks start 136 floats apart, so the 4 blocks a warp sums at
// once fall into different banks.
__device__ __forceinline__ int td_xf_pos(int u) { return u + ((u >> 7) << 3); }

// One tile (clip c, tile index `tile`) in the working precision R of the filter.
template <int NS, typename PCM, typename R>
__device__ __forceinline__ void td_tile(const DevParams& p, const Batch& b, const PCM* __restrict__ pcm,
                                        const int64_t* __restrict__ tile_off, const TdTables& tb, const TdOut& o,
                                        const int c, const int tile, unsigned char* smem_raw) {
    constexpr int DIM = 2 * NS;
    R* s_vend = reinterpret_cast<R*>(smem_raw);                     // [DIM][TD_NT/32] warp-final scan states
    R* s_A = s_vend + DIM * (TD_NT / 32);                           // [DIM*DIM][32] powers of the chunk transition
    R* s_H = s_A + 32 * DIM * DIM;                                  // [TD_CHUNK][DIM]
    double* s_renv = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(s_H + TD_CHUNK * DIM) + 7) & ~(uintptr_t)7);   // [env_cap] raw block envelope (want_block)
    double* s_env = s_renv + tb.env_cap;                            // [env_cap] smoothed block envelope
    float* s_x = reinterpret_cast<float*>(s_env + tb.env_cap);      // [TD_XF] staged PCM, then the float32 result
    __shared__ float s_bsum[2 * TD_FT + 4], s_bmax[2 * TD_FT + 4];   // (block mode at stride 64: two leaves per 128 samples)
    __shared__ int s_near;
    constexpr int L = 256, hop = 128;     // n_fft / hop (enforced by the plan)
    // float32 fast path on int16 input: unscaled samples through the (linear) filter, the scale is applied to the
    // frame statistics
    constexpr bool RAW = sizeof(R) == 4 && sizeof(PCM) == 2;

    const int tid = threadIdx.x;
    const int64_t base = __ldg(b.samp_off + c);
    const int64_t N = __ldg(b.samp_off + c + 1) - base;
    const int64_t f0 = __ldg(b.frame_off + c);
    const int T_clip = (int)(__ldg(b.frame_off + c + 1) - f0);
    const int Tloc = N < L ? 0 : (int)(1 + ((N - L) >> 7));
    const int n_tiles = (int)(__ldg(tile_off + c + 1) - __ldg(tile_off + c));
    const int t0 = tile * TD_FT, t1 = min(Tloc, t0 + TD_FT);
    const bool last = tile == n_tiles - 1;
    const int pad = p.padlen;
    if (tid == 0) s_near = 0;

    // valid x_td range of this tile (clip-relative), then the filter buffer around it
    int64_t vs = (int64_t)t0 * hop - tb.halo; if (vs < 0 || tile == 0) vs = 0;
    int64_t ve = last ? N : (int64_t)(t1 - 1) * hop + L + tb.halo; if (ve > N) ve = N;
    int64_t bs = vs - TD_WARM; if (bs < -pad) bs = -pad;
    int64_t be = ve + TD_WARM; if (be > N + pad) be = N + pad;
    const int len = (int)(be - bs);
    const bool exact_l = bs == -pad, exact_r = be == N + pad;

    APT_STAMP(0);
    // (the callers staged the tables s_A / s_H once per CTA)
    // stage the in-clip part of the buffer as float32 (coalesced 128-bit loads); zeros behind it
    const bool interior = bs >= 0 && be <= N;
    int sh = 0;   // shift of the staged samples that makes the 128-bit staging stores aligned
    if (interior) {
        sh = (4 - (run_head<PCM>(base + bs) & 3)) & 3;
        auto put1 = [&](int i, float v) { s_x[i + sh] = v; };
        auto put4 = [&](int i, float4 v) { *reinterpret_cast<float4*>(s_x + i + sh) = v; };
        load_run<(TD_LB / 4 + TD_NT - 1) / TD_NT + 1, PCM, decltype(put1), decltype(put4), RAW>(pcm, base + bs, len, put1, put4);
    } else {
        for (int i = tid; i < len; i += TD_NT) {
            const int64_t s = bs + i;
            if (s >= 0 && s < N) s_x[i] = load_in<RAW>(pcm, base + s);
        }
    }
    for (int i = len + tid; i < TD_LB; i += TD_NT) s_x[i + sh] = 0.0f;
    __syncthreads();
    APT_STAMP(1);
    // my chunk -> registers (float64).  Every tile but the last of a clip runs all TD_NT chunks full
    // length: the samples behind the buffer end are zeros (a filter at rest stays at rest on zeros, so the
    // backward pass reaches the real data in the same state), which keeps every warp on straight-line
    // code; one partial warp on a predicated path would stall the whole CTA at each barrier.
    // Samples beyond the clip ends are scipy's odd extension 2*x[0]-x[i], 2*x[N-1]-x[N-1-i] (float64).
    const int a0 = tid * TD_CHUNK;
    const int n_mine = exact_r ? max(0, min(TD_CHUNK, len - a0)) : TD_CHUNK;
    R y[TD_CHUNK];
    if (interior) {
#pragma unroll
        for (int j = 0; j < TD_CHUNK; j++) y[j] = (R)s_x[sh + a0 + j];
    } else {
#pragma unroll
        for (int j = 0; j < TD_CHUNK; j++) {
            R v = (R)0;
            if (a0 + j < len) {
                const int64_t s = bs + a0 + j;
                if (s < 0) v = (R)2 * (R)load_in<RAW>(pcm, base) - (R)load_in<RAW>(pcm, base - s);
                else if (s >= N) v = (R)2 * (R)load_in<RAW>(pcm, base + N - 1) - (R)load_in<RAW>(pcm, base + 2 * (N - 1) - s);
                else v = (R)s_x[a0 + j];
            }
            y[j] = v;
        }
    }
    __syncthreads();   // s_x is reused for the result below
    APT_STAMP(2);
    const int ia = exact_r ? (len - 1) / TD_CHUNK : TD_NT - 1;   // last thread holding samples
    iir_pass<NS, false, R>(p, tb, s_A, s_H, y, n_mine, exact_l, y[0], ia, s_vend);
    APT_STAMP(3);
    R xe = (R)0;
    if (exact_r) {
        if (tid == ia) {   // scipy seeds the backward pass with zi * (last forward output)
            // through shared memory (s_x is free between the passes): a run-time index into y[] would
            // move the whole register array to local memory
            R* tmp = reinterpret_cast<R*>(s_x);
#pragma unroll
            for (int j = 0; j < TD_CHUNK; j++) tmp[j] = y[j];
            xe = tmp[n_mine - 1];
        }
    } else if (a0 + TD_CHUNK > len) {   // forward ringing behind the buffer end must not enter the backward pass
#pragma unroll
        for (int j = 0; j < TD_CHUNK; j++) if (a0 + j >= len) y[j] = (R)0;
    }
    iir_pass<NS, true, R>(p, tb, s_A, s_H, y, n_mine, exact_r, xe, ia, s_vend);

    APT_STAMP(4);
    // float32 x_td over the valid range, padded layout
    const int u_off = (int)(bs - (int64_t)t0 * hop) + 128;   // u of buffer index 0
    const int voff = (int)(vs - bs), vend = (int)(ve - bs);
    if (a0 >= voff && a0 + TD_CHUNK <= vend) {
        // whole chunk inside the valid range (almost every thread): the padded position u + 8 * (u / 128) of its 23
        // consecutive samples jumps by 8 at most once, at the 128-sample boundary `bnd` elements in
        const int u0 = a0 + u_off;
        float* pa = s_x + td_xf_pos(u0);
        const int bnd = 128 - (u0 & 127);
#pragma unroll
        for (int j = 0; j < TD_CHUNK; j++) pa[j + (j >= bnd ? 8 : 0)] = to_f32(y[j]);
    } else {
#pragma unroll
        for (int j = 0; j < TD_CHUNK; j++) {
            const int i = a0 + j;
            if (i >= voff && i < vend) s_x[td_xf_pos(i + u_off)] = to_f32(y[j]);
        }
    }
    __syncthreads();
    APT_STAMP(5);
    const int u_vs = voff + u_off;   // u of clip sample vs
    auto xf = [&](int64_t s_clip) -> float { return s_x[td_xf_pos((int)(s_clip - vs) + u_vs)]; };
    if (o.x_td) {
        // each tile owns samples [t0*hop, (t0+TD_FT)*hop), the last tile through N
        int64_t w0 = (int64_t)t0 * hop, w1 = last ? N : (int64_t)(t0 + TD_FT) * hop;
        if (w1 > N) w1 = N;
        for (int64_t s = w0 + tid; s < w1; s += TD_NT) o.x_td[base + s] = xf(s);
    }

    // crest factor (feature_extraction.py:514-523): frame t = 128-sample blocks t and t+1 (hop = 128, L = 256);
    // numpy's pairwise float32 sum of a 256-vector is exactly blocksum(first half) + blocksum(second half),
    // each block summed with 8 strided accumulators, so every block is summed once and shared by two frames.
    const int grp = tid >> 3, lane = tid & 7;
    const unsigned gmask = 0xffu << ((tid & 31) & ~7);
    float* crest_o = o.td + f0;             // (unused in gate mode)
    float* kurt_o = o.td + o.nF + f0;
    const int nfr = max(0, t1 - t0);
    for (int blk = grp; blk < nfr + 1 && nfr > 0; blk += TD_NT / 8) {
        const float* seg = s_x + (blk + 1) * 136;   // u = 128 * (blk + 1)
        float v = seg[lane];
        float r = v * v, pk = fabsf(v);
#pragma unroll
        for (int i = 1; i < 16; i++) { v = seg[8 * i + lane]; r += v * v; pk = fmaxf(pk, fabsf(v)); }
        const float o1 = __shfl_xor_sync(gmask, r, 1);
        const float s1 = (lane & 1) ? o1 + r : r + o1;
        const float o2 = __shfl_xor_sync(gmask, s1, 2);
        const float s2 = (lane & 2) ? o2 + s1 : s1 + o2;
        const float o4 = __shfl_xor_sync(gmask, s2, 4);
        const float bsum = (lane & 4) ? o4 + s2 : s2 + o4;
        pk = fmaxf(pk, __shfl_xor_sync(gmask, pk, 1));
        pk = fmaxf(pk, __shfl_xor_sync(gmask, pk, 2));
        pk = fmaxf(pk, __shfl_xor_sync(gmask, pk, 4));
        if (lane == 0) { s_bsum[blk] = bsum; s_bmax[blk] = pk; }
    }
    __syncthreads();
    if (o.blk_sum) {
        // block mode (geometries other than 256 / 128): the tile hands out the statistics of the 128-sample leaves it owns --
        // those starting in [t0 * 128, (t0 + TD_FT) * 128) at stride g, the clip's last tile through its last leaf -- and
        // crest_blocks_kernel combines them
        const int g = p.td_g, sb = 128 / g;
        // leaves that fit the tile's samples: starts t0 * 128 + j * g with all 128 samples before (t1 + 1) * 128, the clip's
        // last tile up to the clip end (a leaf at an odd multiple of 64 can fit behind the last whole 128-sample block)
        const int nleaf = nfr > 0 ? (last ? (int)((N - (int64_t)t0 * 128 - 128) / g) + 1 : nfr * sb + 1) : 0;
        if (sb > 1) {                                          // stride 64: leaves straddle the padded 128-sample rows
            __syncthreads();
            for (int lf = grp; lf < nleaf; lf += TD_NT / 8) {
                const int u = 128 + lf * g;
                float v = s_x[td_xf_pos(u + lane)];
                float r = v * v, pk = fabsf(v);
#pragma unroll
                for (int i = 1; i < 16; i++) { v = s_x[td_xf_pos(u + 8 * i + lane)]; r += v * v; pk = fmaxf(pk, fabsf(v)); }
                const float o1 = __shfl_xor_sync(gmask, r, 1);
                const float s1 = (lane & 1) ? o1 + r : r + o1;
                const float o2 = __shfl_xor_sync(gmask, s1, 2);
                const float s2 = (lane & 2) ? o2 + s1 : s1 + o2;
                const float o4 = __shfl_xor_sync(gmask, s2, 4);
                const float bsum = (lane & 4) ? o4 + s2 : s2 + o4;
                pk = fmaxf(pk, __shfl_xor_sync(gmask, pk, 1));
                pk = fmaxf(pk, __shfl_xor_sync(gmask, pk, 2));
                pk = fmaxf(pk, __shfl_xor_sync(gmask, pk, 4));
                if (lane == 0) { s_bsum[lf] = bsum; s_bmax[lf] = pk; }
            }
            __syncthreads();
        }
        const int nown = nleaf > 0 ? (last ? nleaf : min(nleaf, TD_FT * sb)) : 0;
        const int64_t bo = td_block_base(base, c, g) + (int64_t)t0 * sb;
        for (int i = tid; i < nown; i += TD_NT) { o.blk_sum[bo + i] = s_bsum[i]; o.blk_max[bo + i] = s_bmax[i]; }
        return;
    }
    if (tid < nfr) {
        const int t = t0 + tid;
        float sumsq = 0.0f + (s_bsum[tid] + s_bsum[tid + 1]);
        float pk = fmaxf(s_bmax[tid], s_bmax[tid + 1]);
        if (RAW) { pk *= 3.0518509447574615e-05f; sumsq *= 3.0518509447574615e-05f * 3.0518509447574615e-05f; }   // 1 / 32767
        const float mean_sq = f_div(sumsq, (float)L);
        const float rms = f_sqrt(mean_sq + d2f(p.eps64));
        const double r = (double)rms;
        float cf = d2f((double)pk / (r > p.eps64 ? r : p.eps64));
        if (isnan(cf) || isinf(cf)) cf = 0.0f;
        if (o.gate) {
            o.gate[f0 + t] = cf > p.gate_thr ? 1 : 0;
            if (o.crest_dbg) o.crest_dbg[f0 + t] = cf;
            if (o.list && fabsf(cf - p.gate_thr) <= o.guard * fabsf(p.gate_thr)) s_near = 1;   // benign race: all write 1
        } else {
            crest_o[t] = cf;
            if (!o.want_kurt) kurt_o[t] = 0.0f;
        }
    }
    APT_STAMP(6);
    if (o.gate) {
        // frames beyond the TD grid have crest factor 0 (zero-fill alignment, rain_frame_classifier.py:178-194)
        if (last)
            for (int t = max(Tloc, 0) + tid; t < T_clip; t += TD_NT) {
                o.gate[f0 + t] = 0.0f > p.gate_thr ? 1 : 0;
                if (o.crest_dbg) o.crest_dbg[f0 + t] = 0.0f;
            }
        __syncthreads();
        if (tid == 0 && o.list && s_near) {
            const int pos = atomicAdd(o.list_count, 1);
            if (pos < o.list_cap) o.list[pos] = make_int2(c, tile);
        }
        return;
    }
    if (o.want_kurt) {
        // unbiased Pearson kurtosis (scipy.stats.kurtosis(fisher=False, bias=False)), numpy float32 sums
        for (int fr = grp; fr < nfr; fr += TD_NT / 8) {
            const int t = t0 + fr;
            const int ub = 128 * (fr + 1);
            auto seg = [&](int i) { return s_x[td_xf_pos(ub + i)]; };
            const float mean = f_div(group8_np_sum([&](int i) { return seg(i); }, L, lane, gmask), (float)L);
            const float m2 = f_div(group8_np_sum([&](int i) { float d = seg(i) - mean; return d * d; }, L, lane, gmask), (float)L);
            const float m4 = f_div(group8_np_sum([&](int i) { double d = (double)(seg(i) - mean); return d2f((d * d) * (d * d)); }, L, lane, gmask), (float)L);
            float kv = 0.0f;
            const float lim = 1.1920929e-07f * mean;
            if (!(m2 <= lim * lim)) {
                const double nn = (double)L;
                float a = f_div(d2f(nn * nn - 1.0) * m4, m2 * m2) - d2f(3.0 * (nn - 1.0) * (nn - 1.0));
                kv = d2f(1.0 / (nn - 2.0) / (nn - 3.0)) * a + 3.0f;
                if (isnan(kv) || isinf(kv)) kv = 0.0f;
            }
            if (lane == 0) kurt_o[t] = kv;
        }
    }
    // frames beyond the TD grid are zero (rain_frame_classifier.py:178-194 zero-fill alignment)
    if (last)
        for (int t = max(Tloc, 0) + tid; t < T_clip; t += TD_NT) {
            crest_o[t] = 0.0f; kurt_o[t] = 0.0f;
            if (o.want_block) { o.td[2 * o.nF + f0 + t] = 0.0f; o.td[3 * o.nF + f0 + t] = 0.0f; o.td[4 * o.nF + f0 + t] = 0.0f; }
        }

    if (o.want_block) {
        // block-RMS envelope (feature_extraction.py:253-366).  Block sums are taken directly in float64
        // (the reference differences a float64 cumulative sum: same value to ~1e-13 relative).
        const int B = max(1, p.blk_len), H = max(1, p.blk_hop);
        const int64_t nb_total = N >= B ? (N - B) / H + 1 : 0;
        const int bstep = max(1, (int)rint((double)hop / (double)H));
        const int bpf = max(1, (L + H - 1) / H);
        const int pp = max(1, p.blk_pp);
        // envelope blocks needed: [bq0, bq1) plus one neighbour each side for the smoother
        int64_t bq0 = (int64_t)t0 * bstep - pp; if (bq0 < 0) bq0 = 0;
        int64_t bq1 = (int64_t)(t1 - 1) * bstep + bpf + pp + 1; if (bq1 > nb_total) bq1 = nb_total;
        const int64_t br0 = bq0 > 0 ? bq0 - 1 : 0, br1 = bq1 < nb_total ? bq1 + 1 : nb_total;
        const int nraw = (int)(br1 - br0);
        double* raw_env = s_renv;
        for (int i = tid; i < nraw; i += TD_NT) {
            const int64_t s = (br0 + i) * H;
            double acc = 0.0;
            for (int e = 0; e < B; e++) { double v = (double)xf(s + e); acc += v * v; }
            const double en = acc / (double)B;
            raw_env[i] = sqrt(en > 0.0 ? en : 0.0);
        }
        __syncthreads();
        const int nenv = (int)(bq1 - bq0);
        for (int i = tid; i < nenv; i += TD_NT) {
            const int64_t bidx = bq0 + i;
            const int r = (int)(bidx - br0);
            double v = raw_env[r];
            if (p.blk_smooth && nb_total >= 3) {
                double acc = 0.0;
                if (bidx > 0) acc = raw_env[r - 1] * 0.25;
                acc = (bidx > 0) ? acc + raw_env[r] * 0.5 : raw_env[r] * 0.5;
                if (bidx + 1 < nb_total) acc += raw_env[r + 1] * 0.25;
                v = acc;
            }
            s_env[i] = v;
        }
        __syncthreads();
        for (int fr = tid; fr < t1 - t0; fr += TD_NT) {
            const int t = t0 + fr;
            int64_t b0 = (int64_t)t * bstep, b1 = b0 + bpf;
            if (b1 > nb_total) b1 = nb_total;
            float oc = 0.0f, ow = 0.0f, orat = 0.0f;
            if (b1 > b0) {
                const int m = (int)(b1 - b0);
                const double* fe = s_env + (b0 - bq0);
                int pi = 0;
                for (int i = 1; i < m; i++) if (fe[i] > fe[pi]) pi = i;
                const double ssum = 0.0 + np_pairwise<double>([&](int i) { return fe[i] * fe[i]; }, 0, m);
                const double rms = sqrt(ssum / (double)m);
                const double pv = fe[pi];
                oc = d2f(pv / (rms > p.eps64 ? rms : p.eps64));
                if (pv > p.eps64 && m >= 3 && pi > 0 && pi < m - 1) {
                    const double lv = fe[pi - 1], rv = fe[pi + 1];
                    if (pv - (lv > rv ? lv : rv) > p.eps64) {
                        const double wv = td_peak_width_half([&](int i) { return fe[i]; }, m, pi);
                        if (isfinite(wv) && wv > 0.0) ow = d2f(wv);
                    }
                }
                const int64_t pidx = b0 + pi;
                const int64_t pre0 = pidx - pp > 0 ? pidx - pp : 0, pre1 = pidx;
                const int64_t po0 = pidx + 1, po1 = pidx + 1 + pp < nb_total ? pidx + 1 + pp : nb_total;
                const double* eg = s_env - bq0;
                double pre = 0.0, post = 0.0;
                if (pre1 > pre0) pre = (0.0 + np_pairwise<double>([&](int i) { return eg[i]; }, (int)pre0, (int)(pre1 - pre0))) / (double)(pre1 - pre0);
                if (po1 > po0) post = (0.0 + np_pairwise<double>([&](int i) { return eg[i]; }, (int)po0, (int)(po1 - po0))) / (double)(po1 - po0);
                orat = d2f(log((post + p.eps64) / (pre + p.eps64)));
                if (isnan(orat) || isinf(orat)) orat = 0.0f;
                if (isnan(oc) || isinf(oc)) oc = 0.0f;
            }
            o.td[2 * o.nF + f0 + t] = oc;
            o.td[3 * o.nF + f0 + t] = ow;
            o.td[4 * o.nF + f0 + t] = orat;
        }
    }
}


// tables of the block-parallel filter -> shared memory, once per CTA (same carve-up as td_tile)
template <int NS, typename R>
__device__ __forceinline__ void td_stage_tables(const TdTables& tb, unsigned char* smem_raw) {
    constexpr int DIM = 2 * NS;
    R* s_A = reinterpret_cast<R*>(smem_raw) + DIM * (TD_NT / 32);
    R* s_H = s_A + 32 * DIM * DIM;
    for (int i = threadIdx.x; i < 32 * DIM * DIM; i += TD_NT) s_A[i] = (R)__ldg(tb.Alin + i);
    for (int i = threadIdx.x; i < TD_CHUNK * DIM; i += TD_NT) s_H[i] = (R)__ldg(tb.H + i);
}

// grid form: blockIdx.y = clip; the CTA walks the tiles b.tile0 + blockIdx.x, + gridDim.x, ... of the launch's range
template <int NS, typename PCM, typename R>
__global__ void __launch_bounds__(TD_NT, (sizeof(R) == 8 ? 2 : 3)) td_features_kernel(const __grid_constant__ DevParams p, Batch b,
                                                               const PCM* __restrict__ pcm,
                                                               const int64_t* __restrict__ tile_off, const __grid_constant__ TdTables tb, TdOut o) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int c = b.clip0 + (int)blockIdx.y;
    const int n_tiles = (int)(__ldg(tile_off + c + 1) - __ldg(tile_off + c));
    const int tile_end = (int)min((int64_t)n_tiles, (int64_t)b.tile0 + (int64_t)b.ntile);
    if (b.tile0 + (int)blockIdx.x >= tile_end) return;
    td_stage_tables<NS, R>(tb, smem_raw);
    for (int tile = b.tile0 + (int)blockIdx.x; tile < tile_end; tile += (int)gridDim.x) {
        td_tile<NS, PCM, R>(p, b, pcm, tile_off, tb, o, c, tile, smem_raw);
        __syncthreads();          // shared memory is reused by the next tile
    }
}

// list form (exact re-check of the float32 fast path): persistent CTAs walk the flagged (clip, tile) pairs
template <int NS, typename PCM>
__global__ void __launch_bounds__(TD_NT, 2) td_recheck_kernel(const __grid_constant__ DevParams p, Batch b,
                                                              const PCM* __restrict__ pcm,
                                                              const int64_t* __restrict__ tile_off, const __grid_constant__ TdTables tb, TdOut o) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = min(*o.list_count, o.list_cap);
    if ((int)blockIdx.x >= n) return;
    TdOut oo = o;
    oo.list = nullptr;            // the exact pass only rewrites gate bytes
    oo.crest_dbg = nullptr;       // (the diagnostic plane keeps the float32 values)
    td_stage_tables<NS, double>(tb, smem_raw);
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const int2 ct = o.list[i];
        td_tile<NS, PCM, double>(p, b, pcm, tile_off, tb, oo, ct.x, ct.y, smem_raw);
        __syncthreads();          // shared memory is reused by the next tile
    }
}

// ---------------------------------------------------------------------------------------------
// Float32 fast path of the TD gate for the default prefilter (two biquad sections), written for instruction count: the
// decision consumes one bit per frame (crest > td_gate_threshold), so nothing here has to match the float64 filter bit
// for bit -- only to stay far inside the guard band (measured deviation of the crest factor ~1e-5 relative, guard 1e-3).
//   tile    = the 56 frames of one tile of the exact kernel (so a flagged tile is re-decided by td_recheck_kernel):
//             7296 samples + 384 / 512 samples of warm-up before / behind = 8192 = 256 threads x 32 samples
//   thread  = 32 consecutive samples in registers through both filter directions (block-parallel IIR as in the exact
//             kernel: direct pass from rest, scan of the chunk-final states, homogeneous response added)
//   crest   = per-thread sum of squares / peak of its 32 samples, 4 threads per 128-sample block by shuffles, two
//             blocks per frame: the filtered waveform never goes back to shared memory
//   edges   = tiles whose buffer would cross a clip end (scipy's odd extension, zi initial state, the zero-filled tail
//             frames) are not computed here at all: they are put on the re-check list
// int16 input runs unscaled through the (linear) filter; the 1/32767 scale is applied to the frame statistics.

// Crest factor of the frames of ANY frame size L = 128 * 2^k at a hop that is a multiple of 128, from the block statistics
// the TD kernel wrote in block mode (feature_extraction.py:514-523).  numpy sums the L float32 squares pairwise:
// vectors longer than 128 are halved down to 128-element leaves, each summed with 8 strided accumulators -- the leaves
// are exactly the 128-sample blocks, and the tree over them is the balanced binary tree in index order.
struct CrestIO {
    const float* blk_sum; const float* blk_max;
    float* td;        // [5][nF]: row 0 crest factor; kurtosis and the block-energy rows are not produced at these geometries
    int64_t nF;
    int mark_missing; // the plane goes to the caller: rows 1..4 are filled with NaN (not computed), not with zeros
};
__global__ void __launch_bounds__(256) crest_blocks_kernel(const __grid_constant__ DevParams p, Batch b, CrestIO io) {
    const int c = b.clip0 + (int)blockIdx.y;
    const int64_t base = __ldg(b.samp_off + c);
    const int64_t N = __ldg(b.samp_off + c + 1) - base;
    const int64_t f0 = __ldg(b.frame_off + c);
    const int T = (int)(__ldg(b.frame_off + c + 1) - f0);
    const int t = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (t >= T) return;
    const int L = p.n_fft, nbk = L >> 7, g = p.td_g, sb = 128 / g, hb = p.hop / g;
    const int Tloc = N < L ? 0 : (int)(1 + (N - L) / p.hop);
    float cf = 0.0f;   // frames beyond the TD grid are zero (rain_frame_classifier.py:178-194)
    if (t < Tloc) {
        const int64_t bo = td_block_base(base, c, g) + (int64_t)t * hb;      // leaf i of the frame starts at t * hop + 128 * i
        float v[32];
        float pk = 0.0f;
        for (int i = 0; i < nbk; i++) { v[i] = __ldg(io.blk_sum + bo + i * sb); pk = fmaxf(pk, __ldg(io.blk_max + bo + i * sb)); }
        for (int w = 1; w < nbk; w <<= 1)
            for (int i = 0; i < nbk; i += 2 * w) v[i] = v[i] + v[i + w];
        const float sumsq = 0.0f + v[0];
        const float mean_sq = f_div(sumsq, (float)L);
        const float rms = f_sqrt(mean_sq + d2f(p.eps64));
        const double r = (double)rms;
        cf = d2f((double)pk / (r > p.eps64 ? r : p.eps64));
        if (isnan(cf) || isinf(cf)) cf = 0.0f;
    }
    io.td[f0 + t] = cf;
    if (io.mark_missing) {
        const float nanv = u2f(0x7fc00000u);
        for (int r = 1; r < APT_N_TD_FEATURES; r++) io.td[(int64_t)r * io.nF + f0 + t] = nanv;
    } else {
        io.td[io.nF + f0 + t] = 0.0f;
    }
}

// ---------------------------------------------------------------------------------------------
constexpr int TDF_NT = 256;
constexpr int TDF_CH = 32;                 // samples per thread
constexpr int TDF_WARM = 384;              // warm-up samples before the first frame (pole radius 0.928: 0.928^384 ~ 3e-13);
                                           // 12 thread chunks = 3 blocks of 128, so frames start on 4-thread groups
constexpr int TDF_LB = TDF_NT * TDF_CH;    // 8192: 384 + 7296 payload + 512 behind
static_assert(TDF_LB - TDF_WARM - ((TD_FT - 1) * 128 + 256) >= 384 && TDF_WARM % 128 == 0, "fast TD tile geometry");
constexpr int TDF_XS = TDF_NT * (TDF_CH + 4) + 8;   // staged samples: 4 floats of padding per thread chunk (conflict-free 128-bit reads)
// float tables (global, built at plan time for chunk = 32): H[32][4], A^(2^k)[5][16], A^pos[32][20 (16 used)]
constexpr int TDF_TAB_H = 0, TDF_TAB_A2K = 128, TDF_TAB_APOS = 208, TDF_TAB_N = 208 + 32 * 20;
struct TdFastParams {
    float c[2][6];        // biquad coefficients
    float thr, guard;     // td_gate_threshold, relative guard band
    float eps;            // feature_extraction eps (1e-9)
    const float* tab;     // [TDF_TAB_N]
    uint8_t* gate;        // [nF]
    float* crest_dbg;     // optional [nF]
    int2* list; int* list_count; int list_cap;
};
inline size_t tdf_smem_bytes() { return sizeof(float) * (TDF_XS + TDF_TAB_N + 4 * (TDF_NT / 32) + 3 * 64); }

template <bool REV>
__device__ __forceinline__ void tdf_pass(const TdFastParams& q, const float* __restrict__ s_tab, float* __restrict__ s_vend,
                                         float (&y)[TDF_CH]) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    constexpr int NW = TDF_NT / 32;
    const float b10 = q.c[0][0], b11 = q.c[0][1], b12 = q.c[0][2], a11 = q.c[0][4], a12 = q.c[0][5];
    const float b20 = q.c[1][0], b21 = q.c[1][1], b22 = q.c[1][2], a21 = q.c[1][4], a22 = q.c[1][5];
    // (1) direct pass from rest
    float z10 = 0.0f, z11 = 0.0f, z20 = 0.0f, z21 = 0.0f;
#pragma unroll
    for (int jj = 0; jj < TDF_CH; jj++) {
        const int j = REV ? TDF_CH - 1 - jj : jj;
        const float x = y[j];
        const float y1 = fmaf(b10, x, z10);
        z10 = fmaf(-a11, y1, fmaf(b11, x, z11));
        z11 = fmaf(-a12, y1, b12 * x);
        const float y2 = fmaf(b20, y1, z20);
        z20 = fmaf(-a21, y2, fmaf(b21, y1, z21));
        z21 = fmaf(-a22, y2, b22 * y1);
        y[j] = y2;
    }
    // (2) scan of the chunk-final states over the chunk order (Kogge-Stone inside the warp, previous warp by one exchange)
    float v0 = z10, v1 = z11, v2 = z20, v3 = z21;
    const int pos = REV ? 31 - lane : lane;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const int d = 1 << k;
        const float u0 = REV ? __shfl_down_sync(0xffffffffu, v0, d) : __shfl_up_sync(0xffffffffu, v0, d);
        const float u1 = REV ? __shfl_down_sync(0xffffffffu, v1, d) : __shfl_up_sync(0xffffffffu, v1, d);
        const float u2 = REV ? __shfl_down_sync(0xffffffffu, v2, d) : __shfl_up_sync(0xffffffffu, v2, d);
        const float u3 = REV ? __shfl_down_sync(0xffffffffu, v3, d) : __shfl_up_sync(0xffffffffu, v3, d);
        if (pos >= d) {
            const float4* A = reinterpret_cast<const float4*>(s_tab + TDF_TAB_A2K + 16 * k);
            const float4 r0 = A[0], r1 = A[1], r2 = A[2], r3 = A[3];
            v0 = fmaf(r0.x, u0, fmaf(r0.y, u1, fmaf(r0.z, u2, fmaf(r0.w, u3, v0))));
            v1 = fmaf(r1.x, u0, fmaf(r1.y, u1, fmaf(r1.z, u2, fmaf(r1.w, u3, v1))));
            v2 = fmaf(r2.x, u0, fmaf(r2.y, u1, fmaf(r2.z, u2, fmaf(r2.w, u3, v2))));
            v3 = fmaf(r3.x, u0, fmaf(r3.y, u1, fmaf(r3.z, u2, fmaf(r3.w, u3, v3))));
        }
    }
    if (pos == 31) { s_vend[4 * w + 0] = v0; s_vend[4 * w + 1] = v1; s_vend[4 * w + 2] = v2; s_vend[4 * w + 3] = v3; }
    // state entering this chunk = scanned state of the previous chunk in order
    float s0 = REV ? __shfl_down_sync(0xffffffffu, v0, 1) : __shfl_up_sync(0xffffffffu, v0, 1);
    float s1 = REV ? __shfl_down_sync(0xffffffffu, v1, 1) : __shfl_up_sync(0xffffffffu, v1, 1);
    float s2 = REV ? __shfl_down_sync(0xffffffffu, v2, 1) : __shfl_up_sync(0xffffffffu, v2, 1);
    float s3 = REV ? __shfl_down_sync(0xffffffffu, v3, 1) : __shfl_up_sync(0xffffffffu, v3, 1);
    if (pos == 0) { s0 = 0.0f; s1 = 0.0f; s2 = 0.0f; s3 = 0.0f; }
    __syncthreads();
    const bool has_prev = REV ? (w < NW - 1) : (w > 0);
    if (has_prev) {
        const float* ve = s_vend + 4 * (REV ? w + 1 : w - 1);
        const float e0 = ve[0], e1 = ve[1], e2 = ve[2], e3 = ve[3];
        const float4* A = reinterpret_cast<const float4*>(s_tab + TDF_TAB_APOS + 20 * pos);    // A^pos
        const float4 r0 = A[0], r1 = A[1], r2 = A[2], r3 = A[3];
        s0 = fmaf(r0.x, e0, fmaf(r0.y, e1, fmaf(r0.z, e2, fmaf(r0.w, e3, s0))));
        s1 = fmaf(r1.x, e0, fmaf(r1.y, e1, fmaf(r1.z, e2, fmaf(r1.w, e3, s1))));
        s2 = fmaf(r2.x, e0, fmaf(r2.y, e1, fmaf(r2.z, e2, fmaf(r2.w, e3, s2))));
        s3 = fmaf(r3.x, e0, fmaf(r3.y, e1, fmaf(r3.z, e2, fmaf(r3.w, e3, s3))));
    }
    // (3) homogeneous response to the entering state
    const float4* H = reinterpret_cast<const float4*>(s_tab + TDF_TAB_H);
#pragma unroll
    for (int m = 0; m < TDF_CH; m++) {
        const int j = REV ? TDF_CH - 1 - m : m;
        const float4 h = H[m];
        y[j] = fmaf(h.x, s0, fmaf(h.y, s1, fmaf(h.z, s2, fmaf(h.w, s3, y[j]))));
    }
    __syncthreads();     // s_vend is reused by the next pass
}

template <typename PCM>
__global__ void __launch_bounds__(TDF_NT, 3) td_gate_fast_kernel(Batch b, const PCM* __restrict__ pcm,
                                                                 const int64_t* __restrict__ tile_off, const TdFastParams q) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_x = reinterpret_cast<float*>(smem_raw);          // [TDF_XS] staged samples
    float* s_tab = s_x + TDF_XS;                              // [TDF_TAB_N]
    float* s_vend = s_tab + TDF_TAB_N;                        // [NW][4]
    float* s_bsum = s_vend + 4 * (TDF_NT / 32);               // [64] per 128-sample block: sum of squares
    float* s_bmax = s_bsum + 64;                              // [64] peak
    float* s_bin = s_bmax + 64;                               // [64] peak of the unfiltered input
    constexpr bool RAW = sizeof(PCM) == 2;
    const int tid = threadIdx.x;
    int64_t tile_in_clip;
    int c;
    if (!tile_clip(b, tile_off, c, tile_in_clip)) return;
    const int tile = (int)tile_in_clip;
    const int64_t base = __ldg(b.samp_off + c);
    const int64_t N = __ldg(b.samp_off + c + 1) - base;
    const int64_t f0 = __ldg(b.frame_off + c);
    const int t0 = tile * TD_FT;
    const int64_t bs = (int64_t)t0 * 128 - TDF_WARM;
    if (bs < 0 || bs + TDF_LB > N) {          // buffer crosses a clip end: the exact kernel decides this tile
        if (tid == 0) {
            const int pos = atomicAdd(q.list_count, 1);
            if (pos < q.list_cap) q.list[pos] = make_int2(c, tile);
        }
        return;
    }
    for (int i = tid; i < TDF_TAB_N; i += TDF_NT) s_tab[i] = __ldg(q.tab + i);
    // stage: coalesced 128-bit loads -> float, sample u at u + 4 * (u / 32) (+ sh: alignment of the vector stores)
    const int sh = (4 - (run_head<PCM>(base + bs) & 3)) & 3;
    {
        auto xpos = [&](int u) { const int v = u + sh; return v + ((v >> 5) << 2); };
        auto put1 = [&](int i, float v) { s_x[xpos(i)] = v; };
        auto put4 = [&](int i, float4 v) { *reinterpret_cast<float4*>(s_x + xpos(i)) = v; };
        load_run<TDF_LB / 4 / TDF_NT + 1, PCM, decltype(put1), decltype(put4), RAW, TDF_NT>(pcm, base + bs, TDF_LB, put1, put4);
    }
    __syncthreads();
    float y[TDF_CH];
    {
        // thread chunk = samples 32 tid .. 32 tid + 31 = staged positions sh + 32 tid + j
        if (sh == 0) {
            const float4* src = reinterpret_cast<const float4*>(s_x + tid * (TDF_CH + 4));
#pragma unroll
            for (int j = 0; j < TDF_CH / 4; j++) { const float4 v = src[j]; y[4 * j] = v.x; y[4 * j + 1] = v.y; y[4 * j + 2] = v.z; y[4 * j + 3] = v.w; }
        } else {
#pragma unroll
            for (int j = 0; j < TDF_CH; j++) { const int v = tid * TDF_CH + j + sh; y[j] = s_x[v + ((v >> 5) << 2)]; }
        }
    }
    float pin = 0.0f;        // peak of the unfiltered input: scales the rounding-error bound of the float32 filter
#pragma unroll
    for (int j = 0; j < TDF_CH; j++) pin = fmaxf(pin, fabsf(y[j]));
    tdf_pass<false>(q, s_tab, s_vend, y);
    tdf_pass<true>(q, s_tab, s_vend, y);
    // frame statistics: thread -> (sum of squares, peak) of its 32 samples, 4 threads -> one 128-sample block
    float ss = 0.0f, pk = 0.0f;
#pragma unroll
    for (int j = 0; j < TDF_CH; j++) { ss = fmaf(y[j], y[j], ss); pk = fmaxf(pk, fabsf(y[j])); }
    ss += __shfl_xor_sync(0xffffffffu, ss, 1); pk = fmaxf(pk, __shfl_xor_sync(0xffffffffu, pk, 1));
    ss += __shfl_xor_sync(0xffffffffu, ss, 2); pk = fmaxf(pk, __shfl_xor_sync(0xffffffffu, pk, 2));
    pin = fmaxf(pin, __shfl_xor_sync(0xffffffffu, pin, 1));
    pin = fmaxf(pin, __shfl_xor_sync(0xffffffffu, pin, 2));
    if ((tid & 3) == 0) { s_bsum[tid >> 2] = ss; s_bmax[tid >> 2] = pk; s_bin[tid >> 2] = pin; }     // block tid/4 of the buffer (64 blocks)
    __syncthreads();
    // frame f = 128-sample blocks TDF_WARM / 128 + f and + f + 1 of the buffer
    bool near = false;
    if (tid < TD_FT) {
        constexpr int B0 = TDF_WARM / 128;
        float sumsq = s_bsum[B0 + tid] + s_bsum[B0 + tid + 1];
        float pkf = fmaxf(s_bmax[B0 + tid], s_bmax[B0 + tid + 1]);
        if (RAW) { pkf *= 3.0518509447574615e-05f; sumsq *= 3.0518509447574615e-05f * 3.0518509447574615e-05f; }   // 1 / 32767
        const float rms = sqrtf(sumsq * (1.0f / 256.0f) + q.eps);
        float cf = pkf / fmaxf(rms, q.eps);
        if (!(cf == cf) || isinf(cf)) cf = 0.0f;
        const int64_t g = f0 + t0 + tid;
        q.gate[g] = cf > q.thr ? 1 : 0;
        if (q.crest_dbg) q.crest_dbg[g] = cf;
        // Guard band: a fixed relative part plus a bound on what float32 rounding in the filter can do to the crest factor.
        // Every operation of the recursion rounds at 2^-24 of values of the size of the INPUT (the states of a high-pass
        // carry the low-frequency content it removes), and the filter spreads an error over its memory: |d crest| / crest
        // <= 2 * C * 2^-24 * peak_in / rms_out with C ~ 10 measured (deviation 5e-6 at peak_in / rms_out ~ 5); C = 100 here,
        // so input dominated by rumble far below the pass band widens the band by itself instead of escaping it.
        float pin_f = 0.0f;
#pragma unroll
        for (int k = -2; k <= 3; k++) pin_f = fmaxf(pin_f, s_bin[B0 + tid + k]);
        if (RAW) pin_f *= 3.0518509447574615e-05f;
        const float tol = q.guard + 200.0f * 5.9604645e-08f * pin_f / fmaxf(rms, 1e-20f);
        near = fabsf(cf - q.thr) <= tol * fabsf(q.thr);
    }
    if (__syncthreads_or(near ? 1 : 0) && tid == 0) {
        const int pos = atomicAdd(q.list_count, 1);
        if (pos < q.list_cap) q.list[pos] = make_int2(c, tile);
    }
}

// ---------------------------------------------------------------------------------------------
// K4+K5+K8+K9: the per-clip time recursions.
//
// Everything that is sequential in time is kept in three SERIAL kernels whose instruction streams
// hold only the loop-carried arithmetic (their run time is chain latency x frames, so every extra
// instruction costs wall time), and everything element-wise around them runs in fully PARALLEL
// kernels; the planes in between travel through HBM/L2:
//   trk1_kernel   serial    lane = (clip, mode bin): noise-PSD tracker pass 1 -> lagged, clamped noise NL
//   flux_kernel   parallel  (clip, frame tile): dB normalisation, positive t-vs-(t-2) flux, per-mode sums
//   base_kernel   serial    lane = (clip, row): float64 quantile baselines -> normalised flux (in place)
//   decide_kernel parallel  frame: TD gate, fixed-band decision, labels and confidences
//   compact_kernel          warp per clip: ordered event indices + counts
//   trk2_kernel   serial    lane = (clip, bin): tracker pass 2 gated by the labels -> noise PSD N2
//   db_kernel     parallel  flat: noise-floor dB plane, its sums, level-0 histogram of the median select
// ---------------------------------------------------------------------------------------------
// trk1 / base address their planes with 32-bit offsets inside a clip (frames x row stride < 2^31, checked by the
// plan): the 64-bit multiply per element costs instructions these loops are made of.  trk2 keeps 64-bit offsets: with
// fewer registers all its CTAs become co-resident and it runs slower.
constexpr int SEQ_KMAX = 128;    // operating-band bins supported (n_fft = 256 -> 71)
#ifndef APT_SEQ_PF
#define APT_SEQ_PF 16
#endif
constexpr int SEQ_PF = APT_SEQ_PF;       // frames per straight-line group of the serial loops (register prefetch)

struct Tracker {
    float trk, ts, nprev;
};

// one step of _update_noise_psd_frame for one bin (rain_signal_processor.py:594-666); t > 0.  Branch-free.
__device__ __forceinline__ float tracker_step(const DevParams& p, Tracker& s, float pk, bool allow, float q, float nq) {
    const float err = pk - s.trk;
    s.ts = p.trk_alpha * s.ts + p.trk_1m_alpha * fabsf(err);
    const float step = p.trk_eta * f_max(s.ts, p.trk_floor);
    const float delta = (pk >= s.trk) ? q * step : nq * step;
    const float cand = f_max(s.trk + delta, 0.0f);
    s.trk = allow ? cand : s.trk;
    const float raw = s.trk;
    const bool up = raw > s.nprev;
    const double lam = up ? p.ema_up : p.ema_down;
    const double oml = up ? (1.0 - p.ema_up) : (1.0 - p.ema_down);
    const double nb = lam * (double)s.nprev + oml * (double)raw;
    // the reference clamps in float64 before rounding: max(min(nb, cap), 0).  cap and 0 are float32
    // values and rounding is monotone, so clamping after the rounding gives the same float32.
    float nf = d2f(nb);
    nf = f_min(nf, p.trk_maxr * pk);     // (cap < nf) ? cap : nf
    nf = f_max(nf, 0.0f);
    s.nprev = nf;
    return nf;
}
__device__ __forceinline__ float tracker_step(const DevParams& p, Tracker& s, float pk, bool allow) {
    return tracker_step(p, s, pk, allow, p.trk_q, p.trk_nq);
}
__device__ __forceinline__ float tracker_first(const DevParams& p, Tracker& s, float pk) {
    s.trk = f_max(pk, 0.0f);
    s.ts = f_max(fabsf(pk), p.trk_floor);
    s.nprev = f_max(f_min(s.trk, p.trk_maxr * pk), 0.0f);
    return s.nprev;
}

// Lane table of pass 1: which band bins are tracked (the mode bins; every band bin when a debug plane
// of pass 1 is requested) and where each mode's lanes start.
struct Trk1Tab {
    int n_lanes;
    int nls;                        // row stride of the NL plane (n_lanes rounded up to 8)
    int mode_l0[APT_MAX_MODES];     // first lane of mode m (its bins are consecutive lanes)
    int mode_n[APT_MAX_MODES];      // bins of mode m
    unsigned char lane_bin[SEQ_KMAX];  // band-relative bin of lane j
    const unsigned short* lane_bin_g;  // the same table in global memory when there are more than SEQ_KMAX lanes, else NULL
    __device__ __forceinline__ int bin_of(int j) const { return lane_bin_g ? (int)__ldg(lane_bin_g + j) : (int)lane_bin[j]; }
};

// Serial-lane bookkeeping shared by the three serial kernels: global lane -> (clip, sub-lane).  Lanes past
// the end of the batch repeat the last real lane's arithmetic with their stores switched off, so that
// every lane of a warp always has a valid clip (no divergence in the main loop).
struct SerialLane {
    int c, sub, T;
    int t_end;          // min(T, b.tb): this lane's frames of the launch are [max(b.ta, first), t_end)
    int emin, emax;     // smallest / largest t_end of the warp (uniform loop bounds)
    int64_t f0;
    int64_t gl;         // global lane index inside the launch's clip range (state slot)
    bool store;
};
__device__ __forceinline__ SerialLane serial_lane(const Batch& b, int per_clip) {
    SerialLane L;
    const int64_t total = (int64_t)b.n_clips * per_clip;
    int64_t gl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    L.store = gl < total;
    if (gl >= total) gl = total - 1;
    L.gl = gl;
    const int ci = (int)(gl / per_clip);
    L.sub = (int)(gl - (int64_t)ci * per_clip);
    L.c = b.clip0 + ci;
    L.f0 = __ldg(b.frame_off + L.c);
    L.T = (int)(__ldg(b.frame_off + L.c + 1) - L.f0);
    L.t_end = min(L.T, b.tb);
    int mn = L.t_end, mx = L.t_end;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    }
    L.emin = mn; L.emax = mx;
    return L;
}

struct Trk1IO {
    const float* P_band;   // [nF][K]
    float* NL;             // [nF][nls] lagged, clamped pass-1 noise of the tracked bins
    float* det_noise_psd;  // optional [nF][K]
    float4* state;         // [plan clips][state_stride]: (trk, ts, nprev, -) carried between time segments
    int state_stride;
    int64_t nF;
};

// Frames [b.ta, b.tb) of every lane; ta == 0 starts the recursion (frame 0: N = P), ta > 0 resumes from `state`.
__global__ void __launch_bounds__(128) trk1_kernel(const __grid_constant__ DevParams p, Batch b,
                                                   const __grid_constant__ Trk1Tab tab, Trk1IO io) {
    const SerialLane L = serial_lane(b, tab.n_lanes);
    const int K = p.K, nls = tab.nls;
    const int kb = tab.bin_of(L.sub);
    const float* Pk = io.P_band + L.f0 * K + kb;
    float* NLk = io.NL + L.f0 * nls + L.sub;
    float* N1k = io.det_noise_psd ? io.det_noise_psd + L.f0 * K + kb : nullptr;
    float4* stp = io.state + (size_t)L.c * io.state_stride + L.sub;
    const int te = L.t_end;
    Tracker tr = {0, 0, 0};
    if (b.ta == 0) {
        // frame 0 (rain_signal_processor.py:700-703): N = P
        const float pk = __ldg(Pk);
        const float n1 = tracker_first(p, tr, pk);
        if (L.store) { NLk[0] = f_min(n1, p.trk_maxr * pk); if (N1k) N1k[0] = n1; }
    } else {
        const float4 st = *stp;
        tr.trk = st.x; tr.ts = st.y; tr.nprev = st.z;
    }
    int t = max(b.ta, 1);
    const int tl = max(te - 1, 0);   // last frame this lane may read
    float pbuf[SEQ_PF];
#pragma unroll
    for (int u = 0; u < SEQ_PF; u++) pbuf[u] = __ldg(Pk + (uint32_t)(min(t + u, tl) * K));
    for (; t + SEQ_PF <= L.emin; t += SEQ_PF) {
        float pc[SEQ_PF];
#pragma unroll
        for (int u = 0; u < SEQ_PF; u++) pc[u] = pbuf[u];
#pragma unroll
        for (int u = 0; u < SEQ_PF; u++) pbuf[u] = __ldg(Pk + (uint32_t)(min(t + SEQ_PF + u, tl) * K));
#pragma unroll
        for (int u = 0; u < SEQ_PF; u++) {
            const float pk = pc[u];
            const float nprev = tr.nprev;                       // N1[t-1]
            const float n1 = tracker_step(p, tr, pk, true);
            const float nl = f_min(nprev, p.trk_maxr * pk);     // lag by one frame, clamp (:874-882)
            if (L.store) { NLk[(uint32_t)((t + u) * nls)] = nl; if (N1k) N1k[(uint32_t)((t + u) * K)] = n1; }
        }
    }
    for (; t < L.emax; t++) {   // ragged tail
        if (t < te) {
            const float pk = __ldg(Pk + (uint32_t)(t * K));
            const float nprev = tr.nprev;
            const float n1 = tracker_step(p, tr, pk, true);
            const float nl = f_min(nprev, p.trk_maxr * pk);
            if (L.store) { NLk[(uint32_t)(t * nls)] = nl; if (N1k) N1k[(uint32_t)(t * K)] = n1; }
        }
    }
    if (L.store) *stp = make_float4(tr.trk, tr.ts, tr.nprev, 0.0f);
}

// ---------------------------------------------------------------------------------------------
// Optional smoothing around the tracker passes (off by default; rain_signal_processor.py:366-396, :690-692, :717-719)
// ---------------------------------------------------------------------------------------------
constexpr int PS_STATE = APT_MAX_PRE_SMOOTH + 1;    // floats of carried state per lane: running sum + ring of past sums

// pre_smooth_frames = L: moving average of the band power over the last L frames from a float32 cumulative sum
// (np.cumsum is sequential in time).  lane = (clip, band bin), frames [b.ta, b.tb), state carried between segments.
__global__ void __launch_bounds__(128) presmooth_kernel(const __grid_constant__ DevParams p, Batch b, const float* __restrict__ P_band,
                                                        float* __restrict__ Psm, float* __restrict__ state, int state_stride) {
    const int K = p.K, Lw = p.pre_smooth;
    const SerialLane L = serial_lane(b, K);
    const float* Pk = P_band + L.f0 * K + L.sub;
    float* Yk = Psm + L.f0 * K + L.sub;
    float* stp = state + ((size_t)L.c * state_stride + L.sub) * PS_STATE;
    float ring[APT_MAX_PRE_SMOOTH];
    float cs = 0.0f;
    if (b.ta > 0) {
        cs = stp[0];
        for (int i = 0; i < Lw; i++) ring[i] = stp[1 + i];
    }
    for (int t = b.ta; t < L.t_end; t++) {
        const float pk = __ldg(Pk + (size_t)t * K);
        const int slot = t % Lw;
        const float old = ring[slot];                       // csum[t - L] once t >= L
        cs = (t == 0) ? pk : cs + pk;
        const float y = (t < Lw) ? f_div(cs, (float)(t + 1)) : f_div(cs - old, (float)Lw);
        ring[slot] = cs;
        if (L.store) Yk[(size_t)t * K] = y;
    }
    if (L.store) {
        stp[0] = cs;
        for (int i = 0; i < Lw; i++) stp[1 + i] = ring[i];
    }
}

// median_frames = L (made odd): Y[t] = np.median(X[max(0, t-L+1) .. t]) per column; even counts (clip start) average the two
// middle values in float32.  Columns = the lanes of `tab` (mode bins or every band bin), planes [nF][K] indexed by band bin.
// grid.x covers (frames of the launch) x lanes, grid.y = clip.
__global__ void __launch_bounds__(256) median_time_kernel(const __grid_constant__ DevParams p, Batch b, const __grid_constant__ Trk1Tab tab,
                                                          const float* __restrict__ X, float* __restrict__ Y) {
    const int c = b.clip0 + (int)blockIdx.y;
    const int64_t f0 = __ldg(b.frame_off + c);
    const int T = (int)(__ldg(b.frame_off + c + 1) - f0);
    const int nl = tab.n_lanes, K = p.K;
    int Lw = p.median; if ((Lw & 1) == 0) Lw += 1;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int t = b.ta + (int)(idx / nl), j = (int)(idx % nl);
    if (t >= min(T, b.tb)) return;
    const int kb = tab.bin_of(j);
    const int t0 = max(0, t - Lw + 1), n = t - t0 + 1;
    float w[APT_MAX_MEDIAN];
    for (int i = 0; i < n; i++) {                            // insertion sort of the window
        const float v = __ldg(X + (f0 + t0 + i) * K + kb);
        int q = i - 1;
        while (q >= 0 && w[q] > v) { w[q + 1] = w[q]; q--; }
        w[q + 1] = v;
    }
    Y[(f0 + t) * K + kb] = (n & 1) ? w[n >> 1] : f_div(w[(n >> 1) - 1] + w[n >> 1], 2.0f);
}

// Lag by one frame and clamp to the current power (rain_signal_processor.py:874-882) as a pass of its own, for the runs in
// which pass 1's result is post-processed (median) or its input is not the power itself (pre-smoothing):
// NL[t][j] = min(N1[max(t-1, 0)][bin_j], maxr * P[t][bin_j]).
__global__ void __launch_bounds__(256) lag_clamp_kernel(const __grid_constant__ DevParams p, Batch b, const __grid_constant__ Trk1Tab tab,
                                                        const float* __restrict__ N1, const float* __restrict__ P_band,
                                                        float* __restrict__ NL, int nls) {
    const int c = b.clip0 + (int)blockIdx.y;
    const int64_t f0 = __ldg(b.frame_off + c);
    const int T = (int)(__ldg(b.frame_off + c + 1) - f0);
    const int nl = tab.n_lanes, K = p.K;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int t = b.ta + (int)(idx / nl), j = (int)(idx % nl);
    if (t >= min(T, b.tb)) return;
    const int kb = tab.bin_of(j);
    const float n1 = __ldg(N1 + (f0 + max(t - 1, 0)) * K + kb);
    NL[(f0 + t) * nls + j] = f_min(n1, p.trk_maxr * __ldg(P_band + (f0 + t) * K + kb));
}

struct FluxIO {
    const float* P_band;   // [nF][K]
    const float* NL; int nls;
    float* mf; int stride; // [nF][stride]: cols 0..M-1 raw per-mode flux, col M weighted total
    float* det_noise_lag; float* D;   // optional [nF][K]
    float* mode_flux;      // optional [M][nF]
    int64_t nF;
    int ft;                // frames per tile (FLUX_FT, or a smaller power of two when the lanes of a tile would not fit)
};
constexpr int FLUX_FT = 256;  // frames per tile (the per-CTA set-up is a large share of a 64-frame tile's instructions)

inline size_t flux_smem_bytes(int ft, int n_lanes) {
    return sizeof(float) * (size_t)(ft + 2) * (n_lanes + 1);
}
// frames per tile for a plan whose widest lane table has n_lanes entries: the D tile stays below ~190 KB
inline int flux_tile_frames(int n_lanes) {
    int ft = FLUX_FT;
    while (ft > 8 && flux_smem_bytes(ft, n_lanes) > 190 * 1024) ft >>= 1;
    return ft;
}

__global__ void __launch_bounds__(256) flux_kernel(const __grid_constant__ DevParams p, Batch b,
                                                   const int64_t* __restrict__ tile_off,
                                                   const __grid_constant__ Trk1Tab tab, FluxIO io) {
    extern __shared__ __align__(16) float s_D[];   // [(FT+2)][n_lanes+1]
    __shared__ float s_mf[FLUX_FT][APT_MAX_MODES];
    __shared__ float s_ltab[64];
    const int tid = threadIdx.x;
    int64_t tile_in_clip;
    int c;
    if (!tile_clip(b, tile_off, c, tile_in_clip)) return;
    const int64_t f0 = __ldg(b.frame_off + c);
    const int T = (int)(__ldg(b.frame_off + c + 1) - f0);
    const int t0 = (int)tile_in_clip * io.ft, nt = min(io.ft, T - t0);
    const int K = p.K, M = p.M, nl_ = tab.n_lanes, nls = io.nls, ds = nl_ + 1;
    if (tid < 64) s_ltab[tid] = u2f(kSvmlLog10TabDev[tid]);
    __syncthreads();
    // detector input D = dB above the lagged noise (rain_signal_processor.py:859-888) for frames t0-2 .. t0+nt-1.
    // Thread = (row group, lane): the lane's bin is fixed, rows advance by 256/LP, three rows' loads in flight.
    const int r0 = t0 >= 2 ? 0 : 2 - t0;           // first row that exists (frames before the clip start do not)
    const int LP = (nl_ + 31) & ~31;               // lanes padded to whole warps
    const int j = tid % LP, rg = tid / LP, RP = 256 / LP;
    if (LP > 256) {
        // more lanes than threads (large frame sizes): flattened (row, lane) walk
        const int rows = nt + 2 - r0;
        for (int idx = tid; idx < rows * nl_; idx += 256) {
            const int r = r0 + idx / nl_, jj = idx - (r - r0) * nl_;
            const int kb = tab.bin_of(jj);
            const int64_t fr = f0 + t0 - 2 + r;
            const float pk = __ldg(io.P_band + fr * K + kb);
            const float nl = p.use_norm ? __ldg(io.NL + fr * nls + jj) : 0.0f;
            float dval;
            if (p.use_norm) {
                if (p.ratio_db) dval = 10.0f * svml_log10f(f_div(pk, nl + p.eps32) + p.eps32, s_ltab);
                else dval = 10.0f * svml_log10f(pk + p.eps32, s_ltab) - 10.0f * svml_log10f(nl + p.eps32, s_ltab);
            } else {
                dval = 10.0f * svml_log10f(pk + p.eps32, s_ltab);
            }
            s_D[r * ds + jj] = dval;
            if (r >= 2) {
                if (io.D) io.D[fr * K + kb] = dval;
                if (io.det_noise_lag) io.det_noise_lag[fr * K + kb] = nl;
            }
        }
    } else if (j < nl_ && rg < RP) {
        const int kb = tab.bin_of(j);
        const float* Pp = io.P_band + (f0 + t0 - 2) * K + kb;
        const float* Np = io.NL + (f0 + t0 - 2) * nls + j;
        constexpr int U = 3;
        for (int rb = r0 + rg; rb < nt + 2; rb += RP * U) {
            float pk[U], nl[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int r = min(rb + u * RP, nt + 1);
                pk[u] = __ldg(Pp + r * K);          // row offsets inside a tile fit 32 bits
                nl[u] = p.use_norm ? __ldg(Np + r * nls) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int r = rb + u * RP;
                if (r < nt + 2) {
                    float dval;
                    if (p.use_norm) {
                        if (p.ratio_db)
                            dval = 10.0f * svml_log10f(f_div(pk[u], nl[u] + p.eps32) + p.eps32, s_ltab);
                        else
                            dval = 10.0f * svml_log10f(pk[u] + p.eps32, s_ltab) - 10.0f * svml_log10f(nl[u] + p.eps32, s_ltab);
                    } else {
                        dval = 10.0f * svml_log10f(pk[u] + p.eps32, s_ltab);
                    }
                    s_D[r * ds + j] = dval;
                    if (r >= 2) {
                        const int64_t gi = (f0 + t0 - 2 + r) * K + kb;
                        if (io.D) io.D[gi] = dval;
                        if (io.det_noise_lag) io.det_noise_lag[gi] = nl[u];
                    }
                }
            }
        }
    }
    __syncthreads();
    // positive t-vs-(t-2) flux summed per mode in numpy order (rain_frame_classifier.py:721-759)
    for (int m = 0; m < M; m++)
    for (int tt = tid; tt < nt; tt += 256) {       // (mode, frame) without an index division
        const int tg = t0 + tt;
        const int lo = tab.mode_l0[m], n = tab.mode_n[m];
        const float* d2 = s_D + (tt + 2) * ds;
        const float* d0 = s_D + tt * ds;
        auto fx = [&](int l) { const float d = d2[l] - d0[l]; return d > 0.0f ? d : (d != d ? d : 0.0f); };
        float s = 0.0f;
        if (tg >= 2 && n > 0) {
            if (n < 8) {
                float r = -0.0f;
                for (int i = 0; i < n; i++) r += fx(lo + i);
                s = 0.0f + r;
            } else {
                s = 0.0f + np_pairwise<float>(fx, lo, n);
            }
        }
        s_mf[tt][m] = s;
        if (io.mode_flux) io.mode_flux[(int64_t)m * io.nF + f0 + tg] = s;
    }
    __syncthreads();
    for (int tt = tid; tt < nt; tt += 256) {
        const int tg = t0 + tt;
        float* row = io.mf + (f0 + tg) * io.stride;
        double tot = 0.0;
        for (int m = 0; m < M; m++) { row[m] = s_mf[tt][m]; tot += p.mode_w[m] * (double)s_mf[tt][m]; }
        row[M] = (tg >= 2) ? d2f(tot) : 0.0f;   // weighted total accumulated in a Python double (:755-759)
    }
}

// lane = (clip, row): row 0 follows the weighted total flux (column M), row r >= 1 follows mode r-1.
// causal_stochastic_low_quantile_baseline in float64 (rain_frame_classifier.py:31-82) and the
// normalisation max(x - b, 0) / (b + norm_min) (:874-893); the column is overwritten with the result.
__device__ __forceinline__ float baseline_step(const DevParams& p, double& bl_base, double& bl_scale, float xf, float ffloor) {
    const double xt = (double)xf;
    float ob = d2f(bl_base);
    if (isnan(ob) || isinf(ob)) ob = ffloor;
    ob = f_max(ob, ffloor);
    const double err = xt - bl_base;
    bl_scale = p.bl_alpha * bl_scale + (1.0 - p.bl_alpha) * fabs(err);
    const double step = p.bl_eta * (bl_scale > p.bl_floor ? bl_scale : p.bl_floor);
    const double delta = (xt >= bl_base) ? p.bl_q * step : -(1.0 - p.bl_q) * step;
    const double nb = bl_base + delta;
    bl_base = nb > p.bl_floor ? nb : p.bl_floor;
    const float ex = f_max(xf - ob, 0.0f);
    float sc = ex;
    if (p.norm_enable) {
        // 0 / den is 0 for the finite positive denominators that occur here; dividing 1 instead keeps
        // __fdiv_rn off its slow path (a zero numerator is the common case)
        // (0 or NaN) / den ends up 0 after the nan_to_num below either way
        const bool pos = ex > 0.0f;
        const float q = f_div(pos ? ex : 1.0f, ob + p.norm_min);
        sc = pos ? q : 0.0f;
    }
    if (isnan(sc) || isinf(sc)) sc = 0.0f;
    return sc;
}

__global__ void __launch_bounds__(128) base_kernel(const __grid_constant__ DevParams p, Batch b, float* mf, int stride,
                                                   double2* state) {
    const int M = p.M;
    const SerialLane L = serial_lane(b, M + 1);
    const int col = (L.sub == 0) ? M : L.sub - 1;
    // plain loads: the column is rewritten in place by this very kernel (the non-coherent path is for data that stays
    // read-only for the kernel's lifetime)
    float* xs = mf + L.f0 * stride + col;
    double2* stp = state + (size_t)L.c * (APT_MAX_MODES + 1) + L.sub;
    const int te = L.t_end;
    const float ffloor = d2f(p.bl_floor);
    double bl_base, bl_scale;
    if (b.ta == 0) {
        const double x0 = (double)xs[0];
        bl_base = x0 > p.bl_floor ? x0 : p.bl_floor;
        bl_scale = fabs(x0) > p.bl_floor ? fabs(x0) : p.bl_floor;
    } else {
        const double2 st = *stp;
        bl_base = st.x; bl_scale = st.y;
    }
    int t = b.ta;
    const int tl = max(te - 1, 0);
    float xbuf[SEQ_PF];
#pragma unroll
    for (int u = 0; u < SEQ_PF; u++) xbuf[u] = xs[(uint32_t)(min(t + u, tl) * stride)];
    for (; t + SEQ_PF <= L.emin; t += SEQ_PF) {
        float xc[SEQ_PF];
#pragma unroll
        for (int u = 0; u < SEQ_PF; u++) xc[u] = xbuf[u];
#pragma unroll
        for (int u = 0; u < SEQ_PF; u++) xbuf[u] = xs[(uint32_t)(min(t + SEQ_PF + u, tl) * stride)];
#pragma unroll
        for (int u = 0; u < SEQ_PF; u++) {
            const float sc = baseline_step(p, bl_base, bl_scale, xc[u], ffloor);
            if (L.store) xs[(uint32_t)((t + u) * stride)] = sc;
        }
    }
    for (; t < L.emax; t++) {
        if (t < te) {
            const float sc = baseline_step(p, bl_base, bl_scale, xs[(uint32_t)(t * stride)], ffloor);
            if (L.store) xs[(uint32_t)(t * stride)] = sc;
        }
    }
    if (L.store) *stp = make_double2(bl_base, bl_scale);
}

struct DecIO {
    const float* mf; int stride;     // [nF][stride] normalised flux: cols 0..M-1 modes, col M total score
    const float* td;                 // [5][nF] (crest row 0, kurtosis row 1)
    const uint8_t* gate_in;          // [nF] TD gate decided upstream (float32 fast path + exact re-check); NULL: from td
    int8_t* frame_class; float* rain_conf; float* noise_conf;
    float* norm_flux; float* score; uint8_t* gate;   // optional
    int64_t nF;
};

// one thread per frame: grid.x covers the frames [b.ta, b.tb) of a clip, grid.y the clips.  TD gate + fixed-band
// decision + labels (rain_frame_classifier.py:230-284, :914-998)
__global__ void __launch_bounds__(256) decide_kernel(const __grid_constant__ DevParams p, Batch b, DecIO io) {
    const int c = b.clip0 + (int)blockIdx.y;
    const int64_t f0 = __ldg(b.frame_off + c);
    const int T = (int)(__ldg(b.frame_off + c + 1) - f0);
    const int t = b.ta + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (t >= min(T, b.tb)) return;
    const int64_t g = f0 + t;
    const float* row = io.mf + g * io.stride;
    const int M = p.M;
    const float score = __ldg(row + M);
    const float f0v = __ldg(row + 0), f1 = __ldg(row + 1), f2 = __ldg(row + 2), f3 = __ldg(row + 3);
    bool gate;
    if (io.gate_in) gate = __ldg(io.gate_in + g) != 0;
    else {
        gate = __ldg(io.td + g) > p.gate_thr;
        if (p.has_ku) gate = gate && (__ldg(io.td + io.nF + g) <= p.ku);
    }
    const float gs = gate ? 1.0f : 0.0f;
    const float l0 = svml_log1pf(f_max(f0v * gs, 0.0f));
    const float l1 = svml_log1pf(f_max(f1 * gs, 0.0f));
    const float l2 = svml_log1pf(f_max(f2 * gs, 0.0f));
    const float l3 = svml_log1pf(f_max(f3 * gs, 0.0f));
    const int hits = (l1 >= p.thr1) + (l2 >= p.thr2) + (l3 >= p.thr3);
    const bool is_rain = (l0 >= p.thr0) && (hits >= max(1, p.min_support));
    const float rc = is_rain ? 1.0f : 0.0f;
    float nc = 1.0f - rc;
    nc = nc < 0.0f ? 0.0f : (nc > 1.0f ? 1.0f : nc);
    const bool weak = (score * gs) <= p.mf_noise_max;
    int8_t cls = 1;
    if (nc >= p.noise_hi && weak && !is_rain) cls = 0;
    if (is_rain) cls = 2;
    if (p.bypass_cls) { cls = 0; io.frame_class[g] = 0; io.rain_conf[g] = 0.0f; io.noise_conf[g] = 1.0f; }
    else { io.frame_class[g] = cls; io.rain_conf[g] = rc; io.noise_conf[g] = nc; }
    if (io.score) io.score[g] = score;
    if (io.gate) io.gate[g] = gate ? 1 : 0;
    if (io.norm_flux)
        for (int m = 0; m < M; m++) io.norm_flux[(int64_t)m * io.nF + g] = __ldg(row + m);
}

// one warp per clip: ascending clip-local indices of the RAIN frames + their count
__global__ void __launch_bounds__(128) compact_kernel(Batch b, const int8_t* __restrict__ frame_class,
                                                      int32_t* __restrict__ event_idx, int32_t* __restrict__ event_count) {
    const int ci = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (ci >= b.n_clips) return;
    const int c = b.clip0 + ci;
    const int64_t f0 = __ldg(b.frame_off + c);
    const int T = (int)(__ldg(b.frame_off + c + 1) - f0);
    const int8_t* fc = frame_class + f0;
    int32_t* ev = event_idx + f0;
    int cnt = 0;
    constexpr int U = 4;
    for (int tb = 0; tb < T; tb += 32 * U) {
        int8_t v[U];
#pragma unroll
        for (int u = 0; u < U; u++) { const int t = tb + u * 32 + lane; v[u] = t < T ? __ldg(fc + t) : (int8_t)0; }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool rain = v[u] == 2;
            const unsigned m = __ballot_sync(0xffffffffu, rain);
            if (rain) ev[cnt + __popc(m & ((1u << lane) - 1u))] = tb + u * 32 + lane;
            cnt += __popc(m);
        }
    }
    if (lane == 0) event_count[c] = cnt;
}

struct Trk2IO {
    const float* P_band;        // [nF][K]
    const int8_t* frame_class;  // [nF]
    float* N2;                  // [nF][K] noise PSD of pass 2
    float4* state;              // [plan clips][state_stride]: (trk, ts, nprev, warm-up count) between time segments
    int state_stride;
    int64_t nF;
    double* aq_state;           // [plan clips] rain_prev_ema between time segments (adaptive quantile only)
};

// NT = CTA size of the launch (128: the compiler takes ~130 registers, three CTAs fit an SM; 256: see the launch site)
// Frames [b.ta, b.tb) of every lane; ta == 0 starts the recursion, ta > 0 resumes from `state`.
// AQ: adaptive tracker quantile (adaptive_q_enable).  Every lane of a clip carries the clip's scalar recursion
// rain_ema <- alpha * rain_ema + (1 - alpha) * [frame excluded], in float64 like the reference's Python floats.
template <int NT, bool AQ = false>
__global__ void __launch_bounds__(NT) trk2_kernel(const __grid_constant__ DevParams p, Batch b, Trk2IO io) {
    const int K = p.K;
    const SerialLane L = serial_lane(b, K);
    const float* Pk = io.P_band + L.f0 * K + L.sub;
    float* Nk = io.N2 + L.f0 * K + L.sub;
    const int8_t* fc = io.frame_class + L.f0;
    float4* stp = io.state + (size_t)L.c * io.state_stride + L.sub;
    const int te = L.t_end;
    Tracker tr = {0, 0, 0};
    int warm;
    double rain_ema = 0.0;
    const double aq_1m = 1.0 - p.aq_alpha;
    auto aq_q = [&](float& q, float& nq) {       // q_eff of the frame about to be processed (:634-638)
        double qe = p.aq_base - (p.aq_base - p.aq_min) * rain_ema;
        qe = qe < p.aq_min ? p.aq_min : (qe > p.aq_base ? p.aq_base : qe);
        q = d2f(qe); nq = d2f(-(1.0 - qe));
    };
    auto aq_upd = [&](int8_t cls) { rain_ema = p.aq_alpha * rain_ema + aq_1m * (cls != 0 ? 1.0 : 0.0); };   // (:663-664)
    if (b.ta == 0) {
        // frame 0 counts as an update when it is allowed (rain_signal_processor.py:700-703)
        const int8_t c0 = __ldg(fc);
        warm = ((0 < p.warm_need) || (c0 == 0)) ? 1 : 0;
        const float n2 = tracker_first(p, tr, __ldg(Pk));
        if (L.store) Nk[0] = n2;
        if (AQ) aq_upd(c0);
    } else {
        const float4 st = *stp;
        tr.trk = st.x; tr.ts = st.y; tr.nprev = st.z; warm = __float_as_int(st.w);
        if (AQ) rain_ema = io.aq_state[L.c];
    }
    int t = max(b.ta, 1);
    const int tl = max(te - 1, 0);
    float pbuf[SEQ_PF];
    int8_t ebuf[SEQ_PF];
#pragma unroll
    for (int u = 0; u < SEQ_PF; u++) {
        const int ti = min(t + u, tl);
        pbuf[u] = __ldg(Pk + (size_t)ti * K);
        ebuf[u] = __ldg(fc + ti);
    }
    for (; t + SEQ_PF <= L.emin; t += SEQ_PF) {
        float pc[SEQ_PF];
        int8_t ec[SEQ_PF];
#pragma unroll
        for (int u = 0; u < SEQ_PF; u++) { pc[u] = pbuf[u]; ec[u] = ebuf[u]; }
#pragma unroll
        for (int u = 0; u < SEQ_PF; u++) {
            const int ti = min(t + SEQ_PF + u, tl);
            pbuf[u] = __ldg(Pk + (size_t)ti * K);
            ebuf[u] = __ldg(fc + ti);
        }
#pragma unroll
        for (int u = 0; u < SEQ_PF; u++) {
            // tracker pass 2: the quantile tracker only moves on warm-up frames and NOISE frames (:1006-1028)
            const bool allow = (warm < p.warm_need) || (ec[u] == 0);
            float q = p.trk_q, nq = p.trk_nq;
            if (AQ) aq_q(q, nq);
            const float n2 = tracker_step(p, tr, pc[u], allow, q, nq);
            if (AQ) aq_upd(ec[u]);
            warm += allow ? 1 : 0;
            if (L.store) Nk[(size_t)(t + u) * K] = n2;
        }
    }
    for (; t < L.emax; t++) {
        if (t < te) {
            const int8_t cls = __ldg(fc + t);
            const bool allow = (warm < p.warm_need) || (cls == 0);
            float q = p.trk_q, nq = p.trk_nq;
            if (AQ) aq_q(q, nq);
            const float n2 = tracker_step(p, tr, __ldg(Pk + (size_t)t * K), allow, q, nq);
            if (AQ) aq_upd(cls);
            warm += allow ? 1 : 0;
            if (L.store) Nk[(size_t)t * K] = n2;
        }
    }
    if (L.store) *stp = make_float4(tr.trk, tr.ts, tr.nprev, __int_as_float(warm));
    if (AQ && L.store && L.sub == 0) io.aq_state[L.c] = rain_ema;
}

// ---------------------------------------------------------------------------------------------
// Optional peak-structure features (rain_frame_classifier.py:761-843; peak_features_enable, default off;
// debug outputs only).  One thread per frame on the detector input D (dB above the lagged noise):
// scipy.signal.find_peaks(D, prominence=, height=median + min_above_floor) in float64 -- local maxima with
// plateau midpoints (_local_maxima_1d), height filter, prominences over the whole row (_peak_prominences,
// wlen=None) -- then the reference's valid-prominence range in float32, per-mode counts, and the top-P
// gate (strongest peaks: share inside any mode band, primary band among the first top_m).
// ---------------------------------------------------------------------------------------------
struct PeakIO {
    const float* D;            // [nF][K]
    float* ratio; float* gate_score; int32_t* valid_count; int32_t* count_by_mode;   // [nF], [nF], [nF], [M][nF]
    int64_t nF;
};

__global__ void __launch_bounds__(128) peak_kernel(const __grid_constant__ DevParams p, Batch b, PeakIO io) {
    const int c = b.clip0 + (int)blockIdx.y;
    const int64_t f0 = __ldg(b.frame_off + c);
    const int T = (int)(__ldg(b.frame_off + c + 1) - f0);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int K = p.K, M = p.M;
    const int64_t g = f0 + t;
    int cnt_mode[APT_MAX_MODES];
    for (int m = 0; m < M; m++) cnt_mode[m] = 0;
    float ratio_o = 0.0f, gate_o = 0.0f;
    int valid_o = 0;
    if (t >= 1) {
        float xf[SEQ_KMAX];
        const float* row = io.D + g * K;
        for (int k = 0; k < K; k++) xf[k] = row[k];
        // np.median of the float32 row: middle element, or the float32 mean of the two middle ones
        float lo = 0.0f, hi = 0.0f;
        for (int k = 0; k < K; k++) {
            int rank = 0;
            for (int j = 0; j < K; j++) rank += (xf[j] < xf[k]) || (xf[j] == xf[k] && j < k);
            if (rank == (K - 1) / 2) lo = xf[k];
            if (rank == K / 2) hi = xf[k];
        }
        const float med = (K & 1) ? lo : f_div(lo + hi, 2.0f);
        const double hthr = (double)med + p.peak_min_above_floor;
        // peaks surviving the height and prominence filters: index, height (float32), prominence (float32)
        unsigned char pk_i[SEQ_KMAX / 2 + 1];
        float pk_h[SEQ_KMAX / 2 + 1];
        int npk = 0;
        int i = 1;
        const int imax = K - 1;
        while (i < imax) {
            if (xf[i - 1] < xf[i]) {
                int ia = i + 1;
                while (ia < imax && xf[ia] == xf[i]) ia++;
                if (xf[ia] < xf[i]) {
                    const int pk = (i + ia - 1) / 2;                 // plateau midpoint
                    const double xp = (double)xf[pk];
                    if (hthr <= xp) {
                        // prominence over the whole row
                        double lmin = xp, rmin = xp;
                        for (int j = pk; j >= 0 && xf[j] <= xf[pk]; j--) if ((double)xf[j] < lmin) lmin = (double)xf[j];
                        for (int j = pk; j <= K - 1 && xf[j] <= xf[pk]; j++) if ((double)xf[j] < rmin) rmin = (double)xf[j];
                        const double prom = xp - (lmin > rmin ? lmin : rmin);
                        if (p.peak_prom_db <= prom) {
                            const float pf = d2f(prom);
                            if (pf >= p.peak_valid_prom_min && pf <= p.peak_valid_prom_max) {   // valid_prom_mask, float32
                                pk_i[npk] = (unsigned char)pk; pk_h[npk] = xf[pk]; npk++;
                            }
                        }
                    }
                    i = ia;
                    continue;
                }
                i = ia;
                continue;
            }
            i++;
        }
        valid_o = npk;
        for (int q = 0; q < npk; q++) {
            const int kb = pk_i[q];
            for (int m = 0; m < M; m++) if (kb >= p.mode_blo[m] && kb <= p.mode_bhi[m]) cnt_mode[m]++;
        }
        if (npk > 0) {
            // strongest top-P valid peaks, tallest first (ties: higher bin first, as a reversed stable argsort)
            const int nsel = min(p.peak_top_p, npk);
            unsigned long long used = 0ull, used_hi = 0ull;
            int in_any = 0;
            bool primary_ok = false;
            const int top_m = min(p.primary_top_m, nsel);
            for (int s_ = 0; s_ < nsel; s_++) {
                int best = -1;
                for (int q = 0; q < npk; q++) {
                    const bool u = q < 64 ? ((used >> q) & 1ull) : ((used_hi >> (q - 64)) & 1ull);
                    if (u) continue;
                    if (best < 0 || pk_h[q] > pk_h[best] || (pk_h[q] == pk_h[best] && pk_i[q] > pk_i[best])) best = q;
                }
                if (best < 64) used |= 1ull << best; else used_hi |= 1ull << (best - 64);
                const int kb = pk_i[best];
                bool any = false;
                for (int m = 0; m < M; m++) any = any || (kb >= p.mode_blo[m] && kb <= p.mode_bhi[m]);
                in_any += any ? 1 : 0;
                if (s_ < top_m && kb >= p.mode_blo[0] && kb <= p.mode_bhi[0]) primary_ok = true;
            }
            const double ratio = (double)in_any / (double)max(1, nsel);
            ratio_o = d2f(ratio);
            gate_o = (primary_ok && ratio >= p.peak_ratio_min) ? 1.0f : 0.0f;
        }
    }
    io.ratio[g] = ratio_o; io.gate_score[g] = gate_o; io.valid_count[g] = valid_o;
    for (int m = 0; m < M; m++) io.count_by_mode[(int64_t)m * io.nF + g] = cnt_mode[m];
}

// ---------------------------------------------------------------------------------------------
// K10: suppressor gain (rain_signal_processor.py:400-533 _compute_gain, :1028-1091).  Runs only when the
// caller asks for G / S_hat (the reference computes it always and discards it under default flags).
// noise_conf is binary on this path (1 - rain_conf): RAIN frames are "rain-like", all others noise-like.
//   gain_kernel       parallel: N_eff = min(N2[lag], maxr P), ratio, raw gain, clip, frequency smoothing
//                     (np.convolve 'same'), per-frame median of the ratio
//   gain_time_kernel  serial in time, lane = (clip, bin): confidence-dependent temporal smoothing, clip
//   shat_kernel       parallel: S_hat = G * S
// ---------------------------------------------------------------------------------------------
struct GainIO {
    const float* P_band; const float* N2; const int8_t* frame_class;
    float* G;            // [nF][K]
    float* ratio_med;    // optional [nF]
    int64_t nF;
    float* snr_mode; float* snr_gate;   // optional [nF] (SNR gating)
};
constexpr int GAIN_FT = 8;        // frames per inner batch (one warp per frame for the median)
constexpr int GAIN_HALF = APT_MAX_GAIN_TAPS / 2;

__global__ void __launch_bounds__(256) gain_kernel(const __grid_constant__ DevParams p, Batch b,
                                                   const int64_t* __restrict__ tile_off, GainIO io) {
    __shared__ float s_g[GAIN_FT][SEQ_KMAX + 2 * GAIN_HALF];
    __shared__ float s_r[GAIN_FT][SEQ_KMAX];
    __shared__ float s_pn[GAIN_FT][2][SEQ_KMAX];   // SNR gating: power and effective noise of the frame's bins
    __shared__ float s_sg[GAIN_FT];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    int64_t tile_in_clip;
    int c;
    if (!tile_clip(b, tile_off, c, tile_in_clip)) return;
    const int64_t f0 = __ldg(b.frame_off + c);
    const int T = (int)(__ldg(b.frame_off + c + 1) - f0);
    const int t0 = (int)tile_in_clip * FLUX_FT, t1 = min(T, t0 + FLUX_FT);
    const int K = p.K, nt = p.n_gain_taps, h = nt / 2;
    for (int i = tid; i < GAIN_FT * (SEQ_KMAX + 2 * GAIN_HALF); i += 256) (&s_g[0][0])[i] = 0.0f;
    for (int tb = t0; tb < t1; tb += GAIN_FT) {
        __syncthreads();
        for (int idx = tid; idx < GAIN_FT * K; idx += 256) {
            const int tt = idx / K, k = idx - tt * K;
            const int t = tb + tt;
            if (t >= t1) continue;
            const float pk = __ldg(io.P_band + (f0 + t) * K + k);
            const int tl = (p.use_lagged && t > 0) ? t - 1 : t;
            const float n = __ldg(io.N2 + (f0 + tl) * K + k);
            const float neff = f_min(n, p.trk_maxr * pk);
            const float ratio = f_div(neff, pk + p.gain_eps);
            s_r[tt][k] = ratio;
            if (p.snr_gate) { s_pn[tt][0][k] = pk; s_pn[tt][1][k] = neff; continue; }
            const bool rain = __ldg(io.frame_class + f0 + t) == 2;
            const float osub = rain ? p.oversub_rain : p.oversub_noise;
            float g;
            if (p.gain_mode == 0) {
                const float r = f_min(f_max(ratio, 0.0f), 1.0f);
                g = 1.0f - osub * f_sqrt(r);
            } else {
                g = f_div(f_max(pk - osub * neff, 0.0f), pk + p.gain_eps);
            }
            s_g[tt][GAIN_HALF + k] = f_min(f_max(g, p.gain_floor), p.gain_ceil);
        }
        __syncthreads();
        if (p.snr_gate) {
            // frame-level SNR over the selected bins: np.sum(axis=0) adds the rows in bin order (:1064-1067), then the gate
            if (tid < GAIN_FT && tb + tid < t1) {
                float pm = 0.0f, nm = 0.0f;
                bool first = true;
                for (int k = 0; k < K; k++)
                    if ((p.snr_mask[k >> 5] >> (k & 31)) & 1u) {
                        if (first) { pm = s_pn[tid][0][k]; nm = s_pn[tid][1][k]; first = false; }
                        else { pm += s_pn[tid][0][k]; nm += s_pn[tid][1][k]; }
                    }
                const float snr = f_div(pm, nm + p.gain_eps);
                float gate = f_div(snr, snr + p.snr1);
                // snr_gating_power (:1073-1075): np.power on the clipped float32 gate; evaluated in float64 and rounded once
                // (numpy's float32 pow is within an ulp of that)
                if (p.snr_pow != 1.0f) gate = d2f(pow((double)f_min(f_max(gate, 0.0f), 1.0f), (double)p.snr_pow));
                s_sg[tid] = f_min(f_max(gate, 0.0f), 1.0f);
                if (io.snr_mode) { io.snr_mode[f0 + tb + tid] = snr; io.snr_gate[f0 + tb + tid] = s_sg[tid]; }
            }
            __syncthreads();
            for (int idx = tid; idx < GAIN_FT * K; idx += 256) {
                const int tt = idx / K, k = idx - tt * K;
                const int t = tb + tt;
                if (t >= t1) continue;
                const bool rain = __ldg(io.frame_class + f0 + t) == 2;
                float osub = rain ? p.oversub_rain : p.oversub_noise;
                if (p.adaptive_gain) osub = osub * (1.0f - s_sg[tt]);          // (:433-438)
                const float pk = s_pn[tt][0][k], neff = s_pn[tt][1][k];
                float g;
                if (p.gain_mode == 0) {
                    const float r = f_min(f_max(s_r[tt][k], 0.0f), 1.0f);
                    g = 1.0f - osub * f_sqrt(r);
                } else {
                    g = f_div(f_max(pk - osub * neff, 0.0f), pk + p.gain_eps);
                }
                s_g[tt][GAIN_HALF + k] = f_min(f_max(g, p.gain_floor), p.gain_ceil);
            }
            __syncthreads();
        }
        for (int idx = tid; idx < GAIN_FT * K; idx += 256) {
            const int tt = idx / K, k = idx - tt * K;
            const int t = tb + tt;
            if (t >= t1) continue;
            const bool rain = __ldg(io.frame_class + f0 + t) == 2;
            const bool smooth = p.gain_freq_smooth && nt > 1 && (!p.adaptive_gain || !rain);
            float g = s_g[tt][GAIN_HALF + k];
            if (smooth) {
                // np.convolve(g, v, 'same') == correlate with the reversed kernel: sum_i g[k-h+i] * v[nt-1-i]
                float acc = 0.0f;
                for (int i = 0; i < nt; i++) acc += s_g[tt][GAIN_HALF + k - h + i] * p.gain_taps[nt - 1 - i];
                g = acc;
            }
            io.G[(f0 + t) * K + k] = g;
        }
        if (io.ratio_med && w < GAIN_FT && tb + w < t1) {
            // np.median over the band: element(s) of rank (K-1)/2 and K/2 by counting
            const float* r = s_r[w];
            float lo = 0.0f, hi = 0.0f;
            for (int k = lane; k < K; k += 32) {
                const float v = r[k];
                int rank = 0;
                for (int j = 0; j < K; j++) rank += (r[j] < v) || (r[j] == v && j < k);
                if (rank == (K - 1) / 2) lo = v;
                if (rank == K / 2) hi = v;
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) { lo += __shfl_xor_sync(0xffffffffu, lo, d); hi += __shfl_xor_sync(0xffffffffu, hi, d); }
            if (lane == 0) io.ratio_med[f0 + tb + w] = (K & 1) ? lo : f_div(lo + hi, 2.0f);
        }
    }
}

__global__ void __launch_bounds__(128) gain_time_kernel(const __grid_constant__ DevParams p, Batch b,
                                                        const int8_t* __restrict__ frame_class, float* __restrict__ G) {
    const int K = p.K;
    const SerialLane L = serial_lane(b, K);
    float* Gk = G + L.f0 * K + L.sub;
    const int8_t* fc = frame_class + L.f0;
    float gt = 0.0f;
    for (int t = 0; t < L.T; t++) {
        const float gf = Gk[(size_t)t * K];
        if (t == 0) gt = gf;
        else if (p.adaptive_gain) {
            // rain-like frames (noise_conf < 0.7): no temporal smoothing; others: alpha = alpha_base * eff_nc
            gt = (__ldg(fc + t) == 2) ? gf : p.alpha_noise * gt + p.om_noise * gf;
        } else {
            gt = p.alpha_base * gt + p.om_base * gf;
        }
        if (L.store) Gk[(size_t)t * K] = f_min(f_max(gt, p.gain_floor), p.gain_ceil);
    }
}

__global__ void __launch_bounds__(256) shat_kernel(const __grid_constant__ DevParams p, int64_t g0, int64_t g1,
                                                   const float* __restrict__ G, const float* __restrict__ S,
                                                   float* __restrict__ S_hat) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (frame, bin) flat over the launch's frames
    const int64_t n = (g1 - g0) * p.F;
    if (i >= n) return;
    const int64_t fr = g0 + i / p.F;
    const int f = (int)(i % p.F);
    const int k = f - p.band_lo;
    const float g = (k >= 0 && k < p.K) ? G[fr * p.K + k] : 1.0f;
    const int64_t o = (fr * p.F + f) * 2;
    S_hat[o] = g * S[o];
    S_hat[o + 1] = g * S[o + 1];
}

// ---------------------------------------------------------------------------------------------
// Inverse STFT of the gain-weighted spectrum (rain_signal_processor.py:1113-1128, librosa.istft with
// center=True, Hann, length=len(x)): per frame the inverse real FFT (float64, as a 128-point complex
// transform of the re-packed half spectrum), synthesis window, overlap-add in frame order, division by
// the window sum-square where it exceeds float32 tiny, float32 output.  n_fft = 256, hop = 128.
// One CTA per ISTFT_TH output hops; frames h0 .. h0+TH contribute.
// ---------------------------------------------------------------------------------------------
constexpr int ISTFT_TH = 8;
constexpr int ISTFT_NT = 256;

__global__ void __launch_bounds__(ISTFT_NT) istft256_kernel(const __grid_constant__ DevParams p, Batch b,
                                                            const float* __restrict__ S_hat, const double* __restrict__ win,
                                                            const cx<double>* __restrict__ tw256, float* __restrict__ y) {
    constexpr int NF = ISTFT_TH + 1, H = 128;
    __shared__ cx<double> bufA[NF][H];
    __shared__ cx<double> bufB[NF][H];
    const int tid = threadIdx.x;
    const int c = b.clip0 + (int)blockIdx.y;
    const int64_t base = __ldg(b.samp_off + c);
    const int64_t N = __ldg(b.samp_off + c + 1) - base;
    const int64_t f0 = __ldg(b.frame_off + c);
    const int T = (int)(__ldg(b.frame_off + c + 1) - f0);
    const int h0 = blockIdx.x * ISTFT_TH;                       // first output hop of this CTA
    if ((int64_t)h0 * 128 >= N) return;
    // re-pack: Z[k] = E[k] + i O[k], E = (X[k] + conj(X[128-k]))/2, O = (X[k] - conj(X[128-k]))/2 * conj(W256^k);
    // the inverse transform is conj(FFT128(conj(Z)))/128, so conj(Z) is what enters the forward passes
    for (int i = tid; i < NF * H; i += ISTFT_NT) {
        const int f = i >> 7, k = i & 127;
        const int t = h0 + f;
        cx<double> zc = {0.0, 0.0};
        if (t < T) {
            const float* X = S_hat + ((f0 + t) * (int64_t)p.F) * 2;
            const cx<double> xk = {(double)X[2 * k], (double)X[2 * k + 1]};
            const cx<double> xn = {(double)X[2 * (128 - k)], -(double)X[2 * (128 - k) + 1]};   // conj(X[128-k])
            const cx<double> e = {(xk.x + xn.x) * 0.5, (xk.y + xn.y) * 0.5};
            const cx<double> d = {(xk.x - xn.x) * 0.5, (xk.y - xn.y) * 0.5};
            const cx<double> o = cmul(d, cconj(tw256[k]));
            zc = {e.x - o.y, -(e.y + o.x)};                      // conj(E + iO)
        }
        bufA[f][k] = zc;
    }
    __syncthreads();
    cx<double>(*x)[H] = bufA;
    cx<double>(*yb)[H] = bufB;
    for (int q = 1; q < H; q <<= 1) {
        const int tstep = 2 * (H / q);                           // W_{2q}^k = W256^(k * 256 / (2q))
        for (int ii = tid; ii < NF * (H >> 1); ii += ISTFT_NT) {
            const int f = ii >> 6, i = ii & 63;
            const int k = i & (q - 1);
            const int j = ((i - k) << 1) + k;
            const cx<double> u0 = x[f][i];
            const cx<double> xv = x[f][i + (H >> 1)];
            const cx<double> u1 = (k == 0) ? xv : cmul(xv, tw256[k * tstep >> 1]);
            yb[f][j] = cadd(u0, u1);
            yb[f][j + q] = csub(u0, u1);
        }
        __syncthreads();
        cx<double>(*tmp)[H] = x; x = yb; yb = tmp;
    }
    // x[f][n] now holds 128 * conj(z[n]):  frame sample 2n = Re / 128, sample 2n+1 = -Im / 128
    for (int i = tid; i < ISTFT_TH * 128; i += ISTFT_NT) {
        const int64_t n = (int64_t)h0 * 128 + i;                 // output sample
        if (n >= N) continue;
        const int64_t pidx = n + 128;                            // index in the centre-padded signal
        const int tb = (int)(pidx >> 7);                         // frames tb-1 and tb cover it
        double acc = 0.0, wss = 0.0;
#pragma unroll
        for (int dt = -1; dt <= 0; dt++) {
            const int t = tb + dt;
            if (t < 0 || t >= T) continue;
            const int m = (int)(pidx - (int64_t)t * 128);        // sample inside frame t
            const cx<double> v = x[t - h0][m >> 1];
            const double fs = ((m & 1) ? -v.y : v.x) * (1.0 / 128.0);
            const double w = win[m];
            acc += w * fs;
            wss += w * w;
        }
        if (wss > 1.1754943508222875e-38) acc /= wss;
        y[base + n] = d2f(acc);
    }
}

// ---------------------------------------------------------------------------------------------
// Noise-floor statistics of a clip: mean and exact median of d = 10*log10(w), w = N2 + eps, over its T*K values
// (rain_signal_processor.py:1286-1298).  The dB values are never stored.
//
// Median.  d(w) -- the float32 log10 polynomial numpy uses, times 10 -- is monotone non-decreasing in w over every
// float32 in [1e-9, 2^20] EXCEPT inside six tiny intervals at w = 1.5 * 2^k (k = -29, -21, -18, -15, 14, 17; at most 24
// consecutive floats each; exhaustive scan, tests/test_kernel_math_emul.py), and every inversion d(a) > d(b), a < b, has
// both a and b inside one of them.  So the order statistics of d are d of the order statistics of w, unless a selected
// value falls into one of those intervals.  The select therefore runs on the bit patterns of w (positive floats order
// like their bits), with no transcendental per element:
//   dbsum_kernel     per time segment, right behind trk2: histogram of bits 30..19 of w (4096 bins of 1/16 octave, about
//                    3 % of a clip's values in the fullest), and the float64 sum of log2(w) (hardware lg2: 2 ulp, far
//                    inside the 1e-5 tolerance of the mean; one multiplication by 10*log10(2) per chunk)
//   sel_scan0_kernel warp per clip: the bins holding the two middle ranks (numpy's even-count rule) + ranks inside;
//                    clips whose bins touch an exclusion interval, or whose w leaves [1e-9, 2^20), are flagged
//   sel_collect_kernel  second read of the plane: the bit patterns of the values in those bins are appended to the clip's
//                    candidate list (warp-aggregated)
//   sel_final_kernel CTA per clip: radix select of both ranks inside the candidate list (L2-resident)
//   finalize_kernel  d of the two selected values with the exact polynomial
// A flagged clip (exclusion interval, candidate overflow of a constant plane, eps < 1e-9) goes through the 3-level radix
// select on the exact dB keys of the whole plane instead (select_hist / select_scan); those kernels return at once for
// every other clip.
// ---------------------------------------------------------------------------------------------
constexpr int SEL_BINS = 2048;       // histogram words per rank slot; the w histogram uses both slots as 4096 bins
constexpr int SEL_WBINS = 4096;
constexpr int DB_CF = 896;           // frames per chunk of dbsum / collect / select_hist (chunk = DB_CF * K values)
struct SelState {
    uint32_t prefix[2];   // the two selected keys: bit patterns of w, or (fallback) order-preserving keys of d
    int64_t rank[2];      // fallback: rank inside the prefix
    int64_t brank[2];     // rank inside the w bin
    int bin[2];           // w bins holding the two middle ranks
    int cnt;              // candidates appended
    int overflow;         // fallback flag
};
__device__ __forceinline__ uint32_t db_key(float v) {
    const uint32_t u = f2u(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_db(uint32_t k) {
    return u2f((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ int w_bin(float w) { return (int)((f2u(w) >> 19) & (SEL_WBINS - 1)); }   // w >= 0
// exclusion intervals of the monotonicity of d(w), as bit patterns [first, last] (oracle: tests/test_kernel_math_emul.py)
__constant__ uint32_t kDbExcl[6][2] = {{0x313ffff1u, 0x31400007u}, {0x353ffff8u, 0x35400000u}, {0x36bffff8u, 0x36c00000u},
                                       {0x383ffff8u, 0x38400000u}, {0x46bffff8u, 0x46c00000u}, {0x483ffff8u, 0x48400000u}};

// chunk `blockIdx.x + b.tile0` of clip `blockIdx.y`: returns false when the clip has no such chunk
__device__ __forceinline__ bool db_chunk(const Batch& b, int K, const int64_t* __restrict__ chunk_off, int& c, int64_t& chunk,
                                         int64_t& f0, int& ne, int64_t& e0) {
    if (!tile_clip(b, chunk_off, c, chunk)) return false;
    f0 = __ldg(b.frame_off + c);
    const int64_t n = (__ldg(b.frame_off + c + 1) - f0) * (int64_t)K;
    e0 = chunk * (int64_t)DB_CF * K;
    const int64_t e1 = min(n, e0 + (int64_t)DB_CF * K);
    ne = (int)(e1 - e0);
    return ne > 0;
}

__global__ void __launch_bounds__(256) dbsum_kernel(const __grid_constant__ DevParams p, Batch b, const float* __restrict__ N2,
                                                    const int64_t* __restrict__ chunk_off, uint32_t* __restrict__ hist,
                                                    double* __restrict__ chunk_sum) {
    __shared__ uint32_t s_h[SEL_WBINS];
    __shared__ double s_part[8];
    const int tid = threadIdx.x;
    int c, ne;
    int64_t chunk, f0, e0;
    if (!db_chunk(b, p.K, chunk_off, c, chunk, f0, ne, e0)) return;
    for (int i = tid; i < SEL_WBINS; i += 256) s_h[i] = 0;
    __syncthreads();
    const float* src = N2 + f0 * p.K + e0;      // the chunk: 32-bit indices from here on
    double acc = 0.0;
    int run_bin = -1;
    uint32_t run_cnt = 0;
    constexpr int U = 8;
    for (int base = tid; base < ne; base += U * 256) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int i = base + u * 256;
            v[u] = i < ne ? __ldg(src + i) : 1.0f;
        }
        float part = 0.0f;      // 8 values in float32, then float64: |log2| < 64, 8 terms -> error below 1e-5 of one value's ulp budget
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int i = base + u * 256;
            if (i < ne) {
                const float w = v[u] + p.eps32;
                part += __log2f(w);
                const int bin = w_bin(w);
                if (bin == run_bin) run_cnt++;
                else {
                    if (run_cnt) atomicAdd(&s_h[run_bin], run_cnt);
                    run_bin = bin; run_cnt = 1;
                }
            }
        }
        acc += (double)part;
    }
    if (run_cnt) atomicAdd(&s_h[run_bin], run_cnt);
    // fixed-order block sum: shuffle tree inside the warp, then warps in order
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((tid & 31) == 0) s_part[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double sm = 0.0;
        for (int w = 0; w < 8; w++) sm += s_part[w];
        chunk_sum[__ldg(chunk_off + c) + chunk] = sm * 3.010299956639812;    // 10 * log10(2)
    }
    uint32_t* hg = hist + (size_t)c * 2 * SEL_BINS;
    for (int i = tid; i < SEL_WBINS; i += 256) {
        const uint32_t cnt = s_h[i];
        if (cnt) atomicAdd(hg + i, cnt);
    }
}

__global__ void select_init_kernel(Batch b, int K, SelState* st) {
    const int ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= b.n_clips) return;
    const int c = b.clip0 + ci;
    const int64_t n = (b.frame_off[c + 1] - b.frame_off[c]) * (int64_t)K;
    SelState z;
    z.prefix[0] = z.prefix[1] = 0;
    z.rank[0] = (n - 1) / 2;
    z.rank[1] = n / 2;
    z.brank[0] = z.brank[1] = 0;
    z.bin[0] = z.bin[1] = 0;
    z.cnt = 0; z.overflow = 0;
    st[c] = z;
}

// one warp: the bin of h[0..nb) holding `rank` and the count of elements in lower bins
__device__ __forceinline__ int warp_find_rank(const uint32_t* __restrict__ h, int nb, int64_t rank, int64_t& before_out) {
    const int lane = threadIdx.x & 31;
    int64_t cum = 0;
    int found = -1;
    for (int base = 0; base < nb && found < 0; base += 32) {
        const uint32_t v = h[base + lane];
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        const bool hit = (cum + (int64_t)inc) > rank;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) {
            const int l = __ffs(m) - 1;
            const uint32_t before = __shfl_sync(0xffffffffu, inc - v, l);
            found = base + l;
            cum += before;
        } else {
            cum += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    before_out = cum;
    return found;
}

__global__ void sel_scan0_kernel(int clip0, int n_clips, float eps32, SelState* st, const uint32_t* __restrict__ hist) {
    const int ci = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (ci >= n_clips) return;
    const int c = clip0 + ci;
    const uint32_t* h = hist + (size_t)c * 2 * SEL_BINS;
    int64_t b0, b1;
    const int f0 = warp_find_rank(h, SEL_WBINS, st[c].rank[0], b0);
    const int f1 = warp_find_rank(h, SEL_WBINS, st[c].rank[1], b1);
    if (lane == 0) {
        const int q0 = f0 < 0 ? 0 : f0, q1 = f1 < 0 ? 0 : f1;
        st[c].bin[0] = q0; st[c].bin[1] = q1;
        st[c].brank[0] = st[c].rank[0] - b0; st[c].brank[1] = st[c].rank[1] - b1;
        // the selected values lie in [w_lo, w_hi]: exact only if that range is inside [1e-9, 2^20) and clear of the
        // exclusion intervals of d(w)
        const int lo_bin = q0 < q1 ? q0 : q1, hi_bin = q0 < q1 ? q1 : q0;
        const uint32_t w_lo = (uint32_t)lo_bin << 19, w_hi = (((uint32_t)hi_bin + 1u) << 19) - 1u;
        bool bad = !(eps32 >= 1e-9f) || w_lo < 0x3089705fu /* 1e-9f */ || w_hi >= 0x49800000u /* 2^20 */ || f0 < 0 || f1 < 0;
        for (int i = 0; i < 6; i++) bad = bad || (w_lo <= kDbExcl[i][1] && w_hi >= kDbExcl[i][0]);
        if (bad) st[c].overflow = 1;
    }
}

constexpr int SEL_STAGE_W = 1024;    // candidates a WARP stages in shared memory before its one append to the clip's list
__global__ void __launch_bounds__(256) sel_collect_kernel(const __grid_constant__ DevParams p, Batch b, const float* __restrict__ N2,
                                                          const int64_t* __restrict__ chunk_off, SelState* st,
                                                          const int64_t* __restrict__ cand_off, uint32_t* __restrict__ cand) {
    __shared__ uint32_t s_buf[8][SEL_STAGE_W];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    int c, ne;
    int64_t chunk, f0, e0;
    if (!db_chunk(b, p.K, chunk_off, c, chunk, f0, ne, e0)) return;
    if (st[c].overflow) return;
    const int bin0 = st[c].bin[0], bin1 = st[c].bin[1];
    const int64_t co = __ldg(cand_off + c);
    const int cap = (int)(__ldg(cand_off + c + 1) - co);
    uint32_t* dst = cand + co;
    const float* src = N2 + f0 * p.K + e0;
    uint32_t* mine = s_buf[w];
    int wn = 0;                                            // candidates this warp has staged (warp-uniform, no atomics)
    constexpr int U = 8;
    for (int base = tid; base - lane - 32 * w < ne; base += U * 256) {     // warp-uniform trip count
        float v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int i = base + u * 256;
            v[u] = i < ne ? __ldg(src + i) : -1.0f;        // -1: never taken
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t key = f2u(v[u] + p.eps32);
            const uint32_t wb = key >> 19;
            const bool take = v[u] >= 0.0f && ((int)wb == bin0 || (int)wb == bin1);
            const unsigned m = __ballot_sync(0xffffffffu, take);
            if (take) {
                const int pos = wn + __popc(m & ((1u << lane) - 1u));
                if (pos < SEL_STAGE_W) mine[pos] = key;
                else {                                     // staging full (a plane with > 12 % of its values in one bin)
                    const int gp = atomicAdd(&st[c].cnt, 1);
                    if (gp < cap) dst[gp] = key;
                }
            }
            wn += __popc(m);
        }
    }
    __syncwarp();
    const int n = min(wn, SEL_STAGE_W);
    int gb = 0;
    if (lane == 0 && n) gb = atomicAdd(&st[c].cnt, n);    // ONE append per warp
    gb = __shfl_sync(0xffffffffu, gb, 0);
    for (int i = lane; i < n; i += 32)
        if (gb + i < cap) dst[gb + i] = mine[i];
}

// one CTA per clip: both ranks inside the candidate list, 8 key bits per pass (keys = bit patterns of w; the 12 top bits
// are the bin)
__global__ void __launch_bounds__(256) sel_final_kernel(int clip0, SelState* st, const int64_t* __restrict__ cand_off,
                                                        const uint32_t* __restrict__ cand) {
    __shared__ uint32_t s_h[2][256];
    __shared__ uint32_t s_pre[2];
    __shared__ int64_t s_rank[2];
    const int c = clip0 + (int)blockIdx.x;
    const int tid = threadIdx.x;
    if (st[c].overflow) return;
    const int64_t co = __ldg(cand_off + c);
    const int cap = (int)(__ldg(cand_off + c + 1) - co);
    const int m = st[c].cnt;
    if (m > cap) { if (tid == 0) st[c].overflow = 1; return; }
    const uint32_t* src = cand + co;
    const uint32_t bin0 = (uint32_t)st[c].bin[0], bin1 = (uint32_t)st[c].bin[1];
    if (tid == 0) { s_pre[0] = s_pre[1] = 0; s_rank[0] = st[c].brank[0]; s_rank[1] = st[c].brank[1]; }
    // keys inside a bin share bits 31..19; the select runs over bits 18..0 in passes of 8, 8 and 3 bits
    for (int pass = 0; pass < 3; pass++) {
        const int sh = pass == 0 ? 11 : (pass == 1 ? 3 : 0);
        const int nb = pass == 2 ? 3 : 8;
        s_h[0][tid] = 0; s_h[1][tid] = 0;
        __syncthreads();
        const uint32_t pre0 = s_pre[0], pre1 = s_pre[1];
        for (int i = tid; i < m; i += 256) {
            const uint32_t key = src[i];
            const uint32_t low = key & 0x7ffffu;
            const uint32_t hi = pass == 0 ? 0u : low >> (sh + nb);
            const uint32_t dg = (low >> sh) & ((1u << nb) - 1u);
            if ((key >> 19) == bin0 && hi == pre0) atomicAdd(&s_h[0][dg], 1u);
            if ((key >> 19) == bin1 && hi == pre1) atomicAdd(&s_h[1][dg], 1u);
        }
        __syncthreads();
        if (tid < 2) {
            int64_t r = s_rank[tid], cum = 0;
            int dgt = (1 << nb) - 1;
            for (int j = 0; j < (1 << nb); j++) {
                const int64_t nx = cum + s_h[tid][j];
                if (nx > r) { dgt = j; break; }
                cum = nx;
            }
            s_pre[tid] = (s_pre[tid] << nb) | (uint32_t)dgt;
            s_rank[tid] = r - cum;
        }
        __syncthreads();
    }
    if (tid == 0) { st[c].prefix[0] = (bin0 << 19) | s_pre[0]; st[c].prefix[1] = (bin1 << 19) | s_pre[1]; }
}

// Fallback for clips flagged `overflow`: 3-level MSD radix select (11 + 11 + 10 key bits) over the whole plane, both
// ranks at once.  level 0: bits 31..21, level 1: bits 20..10, level 2: bits 9..0 of the elements whose higher bits
// equal the prefix found so far.  A CTA histograms one chunk in shared memory and adds the non-empty bins to the
// clip's global histogram.
__global__ void __launch_bounds__(256) select_hist_kernel(const __grid_constant__ DevParams p, Batch b, const float* __restrict__ N2,
                                                          const int64_t* __restrict__ chunk_off, int level,
                                                          const SelState* __restrict__ st, uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_h[2][SEL_BINS];
    __shared__ float s_ltab[64];
    int c, ne;
    int64_t chunk, f0, e0;
    if (!db_chunk(b, p.K, chunk_off, c, chunk, f0, ne, e0)) return;
    if (!st[c].overflow) return;
    const int sh = level == 0 ? 21 : (level == 1 ? 10 : 0);
    const uint32_t mask = level == 2 ? 1023u : 2047u;
    const int shp = level == 1 ? 21 : 10;
    const uint32_t pre0 = st[c].prefix[0], pre1 = st[c].prefix[1];
    const bool two = pre1 != pre0;
    if (threadIdx.x < 64) s_ltab[threadIdx.x] = u2f(kSvmlLog10TabDev[threadIdx.x]);
    for (int i = threadIdx.x; i < 2 * SEL_BINS; i += blockDim.x) (&s_h[0][0])[i] = 0;
    __syncthreads();
    const float* src = N2 + f0 * p.K + e0;
    for (int i = threadIdx.x; i < ne; i += 256) {
        const uint32_t key = db_key(10.0f * svml_log10f(__ldg(src + i) + p.eps32, s_ltab));
        const uint32_t hi = level == 0 ? 0u : key >> shp;
        const int bin = (int)((key >> sh) & mask);
        if (hi == pre0) atomicAdd(&s_h[0][bin], 1u);
        else if (two && hi == pre1) atomicAdd(&s_h[1][bin], 1u);
    }
    __syncthreads();
    uint32_t* hg = hist + (size_t)c * 2 * SEL_BINS;
    for (int i = threadIdx.x; i < 2 * SEL_BINS; i += blockDim.x) {
        const uint32_t cnt = (&s_h[0][0])[i];
        if (cnt) atomicAdd(hg + i, cnt);
    }
}

// one warp per flagged clip: find the bins holding the two ranks, refine prefix / rank (histograms are cleared by
// the host between levels)
__global__ void select_scan_kernel(int clip0, int n_clips, int level, SelState* st, const uint32_t* __restrict__ hist) {
    const int ci = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (ci >= n_clips) return;
    const int c = clip0 + ci;
    if (!st[c].overflow) return;
    const uint32_t pre0 = st[c].prefix[0], pre1 = st[c].prefix[1];
    const int64_t r0 = st[c].rank[0], r1 = st[c].rank[1];
    const bool shared01 = (level == 0) || (pre0 == pre1);
    const int nb = level == 2 ? 1024 : 2048, bits = level == 2 ? 10 : 11;
    const uint32_t* h0 = hist + ((size_t)c * 2 + 0) * SEL_BINS;
    const uint32_t* h1 = hist + ((size_t)c * 2 + (shared01 ? 0 : 1)) * SEL_BINS;
    int64_t b0, b1;
    const int f0 = warp_find_rank(h0, nb, r0, b0);
    const int f1 = warp_find_rank(h1, nb, r1, b1);
    if (lane == 0) {
        st[c].prefix[0] = (pre0 << bits) | (uint32_t)(f0 < 0 ? 0 : f0);
        st[c].prefix[1] = (pre1 << bits) | (uint32_t)(f1 < 0 ? 0 : f1);
        st[c].rank[0] = r0 - b0;
        st[c].rank[1] = r1 - b1;
    }
}

__global__ void finalize_kernel(const __grid_constant__ DevParams p, Batch b, const SelState* __restrict__ st,
                                const double* __restrict__ chunk_sum, const int64_t* __restrict__ chunk_off,
                                const int32_t* __restrict__ event_count,
                                float* __restrict__ stats, int clip_id_base) {
    const int ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= b.n_clips) return;
    const int c = b.clip0 + ci;
    const int T = (int)(b.frame_off[c + 1] - b.frame_off[c]);
    const int cnt = event_count[c];
    const int minf = p.min_frames < 1 ? 1 : p.min_frames;
    const double frac = T > 0 ? (double)cnt / (double)T : 0.0;
    const double med_conf = cnt > 0 ? 1.0 : 0.0;
    double ab = (double)cnt / (double)(2 * minf);
    ab = ab < 0.0 ? 0.0 : (ab > 1.0 ? 1.0 : ab);
    float* r = stats + (size_t)c * APT_N_CLIP_STATS;
    r[0] = (float)(clip_id_base + c);
    r[1] = (float)cnt;
    r[2] = d2f(frac);
    r[3] = cnt >= minf ? 1.0f : 0.0f;
    r[4] = d2f(med_conf > ab ? med_conf : ab);
    r[5] = d2f(med_conf);
    if (p.suppressor_bypass || T == 0) {
        // noise_psd is all zeros: every dB value is 10*log10(eps)
        float lt[64];
        for (int i = 0; i < 64; i++) lt[i] = u2f(kSvmlLog10TabDev[i]);
        r[6] = r[7] = T > 0 ? 10.0f * svml_log10f(p.eps32, lt) : 0.0f;
    } else {
        double s = 0.0;   // chunk sums in chunk order: independent of the launch geometry
        for (int64_t q = chunk_off[c]; q < chunk_off[c + 1]; q++) s += chunk_sum[q];
        r[6] = d2f(s / ((double)T * (double)p.K));
        float a, bb;
        if (st[c].overflow) { a = key_db(st[c].prefix[0]); bb = key_db(st[c].prefix[1]); }   // fallback: keys of d
        else {   // bit patterns of w: d with the exact polynomial
            float lt[64];
            for (int i = 0; i < 64; i++) lt[i] = u2f(kSvmlLog10TabDev[i]);
            a = 10.0f * svml_log10f(u2f(st[c].prefix[0]), lt);
            bb = 10.0f * svml_log10f(u2f(st[c].prefix[1]), lt);
        }
        r[7] = f_div(a + bb, 2.0f);
    }
}

// exactness self-tests of the branch-free float32 division / square root (apt_selftest)
__global__ void selftest_sqrt_kernel(unsigned long long* bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t u = 0x3f800000u + i; u <= 0x40000000u; u += gridDim.x * blockDim.x) {
        const float v = __uint_as_float(u);
        if (__float_as_uint(f_sqrt_12(v)) != __float_as_uint(__fsqrt_rn(v))) atomicAdd(bad, 1ull);
    }
}
__global__ void selftest_div_kernel(unsigned long long* bad, unsigned long long n) {
    const unsigned long long i0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (unsigned long long i = i0; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        // two hashed operands: magnitudes log-uniform over 2^-60 .. 2^20, full mantissas, a <= b (the cabs
        // use) for even i, arbitrary order (the baseline normalisation) for odd i
        unsigned long long h = i * 0x9E3779B97F4A7C15ull; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
        const uint32_t ma = (uint32_t)h & 0x7fffffu, mb = (uint32_t)(h >> 23) & 0x7fffffu;
        const uint32_t ea = 67u + (uint32_t)((h >> 46) % 80u), eb = 67u + (uint32_t)((h >> 54) % 80u);
        float a = __uint_as_float((ea << 23) | ma), b = __uint_as_float((eb << 23) | mb);
        if (!(i & 1) && a > b) { const float t = a; a = b; b = t; }
        if (fabsf(__log2f(a) - __log2f(b)) > 100.0f) continue;   // quotient outside the exact path's contract
        if (__float_as_uint(f_div_nr(a, b)) != __float_as_uint(__fdiv_rn(a, b))) atomicAdd(bad, 1ull);
    }
}

}  // namespace apt
