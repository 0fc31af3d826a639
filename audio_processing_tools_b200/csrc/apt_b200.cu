// apt_b200.cu -- C ABI (include/apt_b200.h) over the kernels in apt_kernels.cuh.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -shared -Xcompiler -fPIC
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "apt_kernels.cuh"
#include "apt_dsd.cuh"
#include "apt_bne.cuh"
#include "apt_roe.cuh"
#include "apt_tcdft.cuh"

using namespace apt;

// grow-only device scratch of the engines beside the main path (apt_dsd_run_i16 / apt_bne_run / apt_roe_run): one slot
// per buffer, kept from call to call so that a steady stream of same-sized batches never reaches cudaMalloc
struct ScratchPool {
    std::vector<void*> ptr;
    std::vector<size_t> cap;
    cudaError_t get(int slot, size_t bytes, void** out) {
        if ((int)ptr.size() <= slot) { ptr.resize(slot + 1, nullptr); cap.resize(slot + 1, 0); }
        if (cap[slot] < bytes) {
            if (ptr[slot]) cudaFree(ptr[slot]);
            ptr[slot] = nullptr; cap[slot] = 0;
            const size_t want = bytes + bytes / 4 + 256;
            cudaError_t e = cudaMalloc(&ptr[slot], want);
            if (e != cudaSuccess) return e;
            cap[slot] = want;
        }
        *out = ptr[slot];
        return cudaSuccess;
    }
    ~ScratchPool() { for (void* q : ptr) if (q) cudaFree(q); }
};

struct apt_ctx {
    int device;
    int sm_count;
    std::string err;
    ScratchPool pool;
};

static int fail(apt_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}
#define CUDA_OK(ctx, call)                                                                         \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(ctx, -10, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        free_();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void**)&p, count * sizeof(T));
    }
    void free_() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { free_(); }
};
// same interface on a slot of the context's scratch pool (not owned)
template <typename T>
struct PoolBuf {
    T* p = nullptr;
    size_t n = 0;
    apt_ctx* ctx;
    int slot;
    PoolBuf(apt_ctx* c, int s) : ctx(c), slot(s) {}
    cudaError_t alloc(size_t count) {
        n = count;
        void* q = nullptr;
        cudaError_t e = ctx->pool.get(slot, std::max<size_t>(count, 1) * sizeof(T), &q);
        p = (T*)q;
        return e;
    }
};

struct apt_plan {
    apt_ctx* ctx = nullptr;
    apt_params_t prm;
    DevParams dp;
    int n_clips = 0;
    std::vector<int64_t> len, samp_off, frame_off, stft_tile_off, td_tile_off, sel_chunk_off, flux_tile_off;
    int64_t nS = 0, nF = 0;
    DevBuf<int64_t> d_samp_off, d_frame_off, d_stft_tile_off, d_td_tile_off, d_sel_chunk_off, d_flux_tile_off;
    // tables
    DevBuf<double> d_win64; DevBuf<cx<double>> d_tw128_64, d_tw256_64;
    DevBuf<float> d_win32;  DevBuf<cx<float>> d_tw128_32, d_tw256_32;
    DevBuf<float> d_freqs;
    DevBuf<double> d_Apow, d_H;
    TdTables tdt;
    int td_ns = 0;
    size_t td_smem = 0, td_smem_f32 = 0;
    // float32 fast path of the TD gate: gate bytes, flagged tiles of each time segment, counters
    DevBuf<uint8_t> d_gate;
    DevBuf<int2> d_td_list;
    DevBuf<int> d_td_cnt;
    int64_t td_list_cap = 0;
    int td_fast = 1;
    // tensor-core DFT (fft mode 2): B image (window x twiddles in two fp16 limbs, shared-memory layout), error flag
    DevBuf<unsigned char> d_tc_B;
    DevBuf<int> d_tc_err;
    bool tdf_ok = false;                 // the dedicated float32 gate kernel (two biquad sections) is usable
    DevBuf<float> d_tdf_tab;
    TdFastParams tdf;
    float td_guard = 2.5e-4f;
    // scratch
    DevBuf<float> d_Pband, d_n2, d_td, d_mf, d_nl, d_nl_all, d_Dscr;
    DevBuf<double> d_dbsum;   // [dB chunks] float64 sums of the noise-floor dB values
    // state carried between the time segments of the pipelined run
    DevBuf<float4> d_st_trk1, d_st_trk2;   // [clips][K]
    DevBuf<double2> d_st_base;             // [clips][APT_MAX_MODES + 1]
    DevBuf<double> d_st_aq;                // [clips] rain_prev_ema of the adaptive tracker quantile
    // optional smoothing around the tracker passes (pre_smooth_frames / median_frames): smoothed band power and its carried
    // state, raw pass-1 / pass-2 planes ahead of the median, filtered pass-1 plane
    DevBuf<float> d_Psm, d_st_ps, d_N1raw, d_N1med, d_N2raw;
    // candidate lists of the median select
    std::vector<int64_t> cand_off;
    DevBuf<int64_t> d_cand_off;
    DevBuf<uint32_t> d_cand;
    Trk1Tab tab_modes, tab_all;   // pass-1 lane tables: mode bins only / every band bin (debug planes)
    int mf_stride = 8;
    int st_stride = 0;      // lanes per clip in the carried-state arrays
    bool generic = false;   // generic frame-size STFT kernel
    bool full_ok = true;    // the full pipeline is planned (n_fft = 256 / hop = 128, or any supported frame size at a hop that is a multiple of 64)
    bool td_blocks = false; // geometry other than 256 / 128: TD crest factor from block statistics (one segment, no fast gate)
    int flux_ft = FLUX_FT;  // frames per tile of the flux kernel
    DevBuf<unsigned short> d_lane_modes, d_lane_all;   // pass-1 lane tables beyond SEQ_KMAX lanes
    DevBuf<float> d_blk_sum, d_blk_max;                // 128-sample block statistics of the prefiltered waveform (generic geometry)
    DevBuf<double> d_gwin64; DevBuf<cx<double>> d_gtw64; DevBuf<float> d_gwin32; DevBuf<cx<float>> d_gtw32;
    DevBuf<SelState> d_sel;
    DevBuf<uint32_t> d_hist;
    DevBuf<int> d_counter;
    // host-path staging
    DevBuf<int16_t> d_pcm;
    DevBuf<float> d_pcm_f32;
    static constexpr int N_RING = 3;          // pinned slots of the pageable-clips path (apt_run_host_clips)
    void* ring[N_RING] = {nullptr, nullptr, nullptr};
    size_t ring_bytes = 0;
    DevBuf<int8_t> d_fc; DevBuf<float> d_rc, d_nc, d_stats; DevBuf<int32_t> d_ev, d_evc;
    static constexpr int N_COMP = 4;
    cudaStream_t s_copy = nullptr, s_comp[N_COMP] = {nullptr, nullptr, nullptr, nullptr};
    // Pipelined run: one stream per kernel kind, so that kernel k of time segment s+1 overlaps kernel k+1 of segment s;
    // events chain the kinds inside a segment.  The bulk kernels (STFT, TD) run at the lowest priority, everything
    // on the latency-bound chain behind them at the highest.
    enum { SK_STFT, SK_TD, SK_TRK1, SK_FLUX, SK_BASE, SK_DEC, SK_TRK2, SK_DBS, SK_N };
    static constexpr int MAX_SEG = 64;
    cudaStream_t s_kind[SK_N] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_seg[SK_N][MAX_SEG] = {};
    cudaEvent_t ev_fork = nullptr;
    // optional timeline of the pipelined run (apt_plan_trace): timing events around every launch on its kind's stream
    bool trace = false;
    cudaEvent_t tr_origin = nullptr, tr_end = nullptr, tr_ev[SK_N][MAX_SEG][2] = {};
    int tr_nseg = 0;
    int pipeline = 1;                        // 0: everything on the caller's stream in one segment
    int last_launches = 0;
    size_t scratch_bytes = 0;
    // optional per-kernel timing (CUDA events on the launch stream)
    bool timing = false;
    std::vector<std::pair<int, cudaEvent_t>> marks;   // (kernel id, event recorded BEFORE that kernel); id -1 = end
    float kernel_ms[APT_N_KERNELS] = {0};
    void mark(int id, cudaStream_t st) {
        if (!timing) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        marks.emplace_back(id, e);
    }
};

static void build_td_tables(const apt_params_t& prm, int ns, const double sos[][6], int chunk,
                            std::vector<double>& Apow, std::vector<double>& H) {
    const int dim = 2 * ns;
    std::vector<double> A(dim * dim, 0.0);
    H.assign((size_t)chunk * dim, 0.0);
    for (int r = 0; r < dim; r++) {
        std::vector<double> z(dim, 0.0);
        z[r] = 1.0;
        for (int n = 0; n < chunk; n++) {
            double x = 0.0;
            for (int s = 0; s < ns; s++) {
                const double* c = sos[s];
                double y = c[0] * x + z[2 * s];
                z[2 * s] = c[1] * x - c[4] * y + z[2 * s + 1];
                z[2 * s + 1] = c[2] * x - c[5] * y;
                x = y;
            }
            H[(size_t)n * dim + r] = x;
        }
        for (int q = 0; q < dim; q++) A[q * dim + r] = z[q];
    }
    // A^e for e = 0..31, component-major: Alin[(i*dim+j)*32 + e]
    Apow.assign((size_t)32 * dim * dim, 0.0);
    std::vector<double> cur(dim * dim, 0.0), nxt(dim * dim);
    for (int i = 0; i < dim; i++) cur[i * dim + i] = 1.0;
    for (int e = 0; e < 32; e++) {
        for (int i = 0; i < dim * dim; i++) Apow[(size_t)i * 32 + e] = cur[i];
        for (int i = 0; i < dim; i++)
            for (int j = 0; j < dim; j++) {
                double s = 0.0;
                for (int q = 0; q < dim; q++) s += A[i * dim + q] * cur[q * dim + j];
                nxt[i * dim + j] = s;
            }
        cur = nxt;
    }
    (void)prm;
}

template <typename B, typename T>
static cudaError_t upload(B& b, const std::vector<T>& h) {
    cudaError_t e = b.alloc(h.size());
    if (e != cudaSuccess) return e;
    if (h.empty()) return cudaSuccess;
    return cudaMemcpy(b.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

extern "C" {

void apt_plan_destroy(apt_plan_t* plan);

#ifndef APT_SRC_HASH
#define APT_SRC_HASH "unhashed"
#endif
static const char kSrcHashTag[] = "APT_SRC_HASH=" APT_SRC_HASH;   // found by the loader in the file's bytes
const char* apt_source_hash(void) { return kSrcHashTag + 13; }
int apt_abi_version(void) { return APT_ABI_VERSION; }
int apt_sizeof_params(void) { return (int)sizeof(apt_params_t); }
int apt_sizeof_out(void) { return (int)sizeof(apt_out_t); }

int apt_init(int device_ordinal, apt_ctx** out) {
    if (!out) return -1;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return -2;   // no CUDA device: there is no CPU path
    if (device_ordinal < 0 || device_ordinal >= n) return -3;
    apt_ctx* c = new apt_ctx();
    c->device = device_ordinal;
    if (cudaSetDevice(device_ordinal) != cudaSuccess) { delete c; return -4; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess) { delete c; return -5; }
    c->sm_count = prop.multiProcessorCount;
    *out = c;
    return 0;
}

void apt_destroy(apt_ctx* ctx) { delete ctx; }

const char* apt_last_error(apt_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context (is a CUDA device visible?)"; }

int apt_selftest(apt_ctx* ctx, int which, int64_t n, int64_t* mismatches) {
    if (!ctx || !mismatches) return -1;
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    unsigned long long* d = nullptr;
    CUDA_OK(ctx, cudaMalloc((void**)&d, 8));
    CUDA_OK(ctx, cudaMemset(d, 0, 8));
    if (which == 0) selftest_sqrt_kernel<<<ctx->sm_count * 8, 256>>>(d);
    else if (which == 1) selftest_div_kernel<<<ctx->sm_count * 8, 256>>>(d, (unsigned long long)n);
    else { cudaFree(d); return fail(ctx, -1, "apt_selftest: unknown test %d", which); }
    unsigned long long h = 0;
    cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, -10, "apt_selftest: %s", cudaGetErrorString(e));
    *mismatches = (int64_t)h;
    return 0;
}

int apt_params_default(apt_params_t* p) {
    if (!p) return -1;
    memset(p, 0, sizeof(*p));
    p->abi_version = APT_ABI_VERSION;
    p->fs = 11162; p->n_fft = 256; p->hop = 128;
    p->band_lo = 10; p->band_hi = 80;
    p->n_modes = 0;
    for (int i = 0; i < APT_MAX_MODES; i++) { p->mode_lo[i] = 1; p->mode_hi[i] = 0; p->mode_band_lo[i] = 1; p->mode_band_hi[i] = 0; p->mode_weight[i] = 1.0; }
    // W = max(10, int(0.5 * 11162/128)) = 43
    p->trk_eta = (float)(2.0 / 44.0); p->trk_scale_alpha = 0.95f; p->trk_one_minus_alpha = (float)(1.0 - 0.95);
    p->trk_step_floor = 1e-9f; p->trk_q = 0.25f; p->trk_neg_one_minus_q = -0.75f; p->trk_maxr = 1.0f;
    p->ema_up = 0.6; p->ema_down = 0.95; p->warmup_need = 21; p->eps_f32 = 1e-9f;
    p->detector_use_noise_norm = 1; p->norm_ratio_db = 0;
    p->bl_q = 0.2; p->bl_eta = 2.0 / 45.0; p->bl_scale_alpha = 1.0 - 2.0 / 45.0; p->bl_floor = 1.0;
    p->norm_enable = 1; p->norm_min_f32 = 1.0f;
    p->thr_primary = 1.8f; p->thr_m1 = 2.6f; p->thr_m2 = 2.6f; p->thr_m3 = 3.0f; p->min_support = 2;
    p->td_gate_thr = 2.5f; p->has_kurt_upper = 0; p->kurt_upper = 0.0f;
    p->noise_hi = 0.8f; p->mode_flux_noise_max = 1.5f;
    p->n_sos = 0; p->padlen = 0; p->eps_f64 = 1e-9;
    p->blk_len = 8; p->blk_hop = 8; p->blk_post_pre = 4; p->blk_smooth = 1;
    p->low_lo = 2; p->low_hi = 4; p->rain_lo = 10; p->rain_hi = 18; p->rolloff_fraction = 0.85;
    p->suppressor_bypass = 0; p->clip_rain_min_frames = 1; p->fft_f64 = 1;
    p->adaptive_q = 0; p->aq_base = 0.25; p->aq_min = 0.10; p->aq_alpha = 0.95;
    p->gain_mode = 0; p->adaptive_gain = 1; p->gain_freq_smooth = 1; p->n_gain_taps = 3; p->use_lagged_noise_psd = 0;
    p->oversub_noise = 3.0f; p->oversub_rain = 1.0f; p->gain_floor = 0.0f; p->gain_ceil = 1.0f;
    p->gain_taps[0] = 0.2f; p->gain_taps[1] = 0.6f; p->gain_taps[2] = 0.2f;
    p->alpha_noise = 0.7f; p->one_minus_alpha_noise = 0.3f; p->alpha_base = 0.7f; p->one_minus_alpha_base = 0.3f;
    p->gain_eps_f32 = 1e-9f;
    p->peak_top_p = 6; p->primary_top_m = 3; p->peak_prominence_db = 3.0; p->peak_min_db_above_floor = 6.0; p->peak_ratio_min = 0.5;
    p->peak_valid_prom_min_db = 3.0f; p->peak_valid_prom_max_db = 6.0f;
    return 0;
}

int apt_plan_create(apt_ctx* ctx, const apt_params_t* p, int n_clips, const int64_t* clip_len, apt_plan_t** out) {
    if (!ctx) return -1;
    if (!p || !out || !clip_len || n_clips <= 0) return fail(ctx, -1, "apt_plan_create: bad arguments");
    if (n_clips > 65535) return fail(ctx, -1, "apt_plan_create: at most 65535 clips per plan (grid.y limit), got %d", n_clips);
    *out = nullptr;
    if (p->abi_version != APT_ABI_VERSION) return fail(ctx, -20, "params abi_version %d != %d", p->abi_version, APT_ABI_VERSION);
    // (256, 128) runs the full pipeline on the specialised kernels.  Any other power-of-two frame size up to 4096 with
    // 1 <= hop <= n_fft runs the features stage (STFT, power, band energies, raw features) on the generic STFT kernel,
    // and the full pipeline too when the hop is a multiple of 64 (the TD crest factor is then assembled from 128-sample
    // block statistics; the kurtosis gate, the peak features and the gain planes stay with the 256-sample geometry)
    if (p->n_fft < 256 || p->n_fft > 4096 || (p->n_fft & (p->n_fft - 1)) != 0 || p->hop < 1 || p->hop > p->n_fft)
        return fail(ctx, -21, "unsupported STFT geometry n_fft=%d hop=%d (power of two 256..4096, 1 <= hop <= n_fft)", p->n_fft, p->hop);
    // n_fft = 256 with any hop <= 128 stays on the specialised STFT kernel (features stage unless hop = 128)
    const bool generic = !(p->n_fft == 256 && p->hop <= 128);
    // full pipeline: 256 / 128, or any supported frame size at a hop that is a multiple of 64 (128-sample leaves of numpy's
    // pairwise sums start at multiples of the block stride g = 128 or 64)
    const bool td_blocks = !(p->n_fft == 256 && p->hop == 128);
    const bool full_ok = !td_blocks || (p->hop % 64 == 0 && !p->has_kurt_upper);
    const int F = p->n_fft / 2 + 1;
    const int K = p->band_hi - p->band_lo + 1;
    if (p->band_lo < 0 || p->band_hi >= F || K < 1 || (!generic && K > SEQ_KMAX)) return fail(ctx, -22, "operating band bins [%d,%d] unsupported", p->band_lo, p->band_hi);
    if (p->n_modes < 4 || p->n_modes > APT_MAX_MODES) return fail(ctx, -23, "n_modes=%d outside [4,%d]", p->n_modes, APT_MAX_MODES);
    if (p->n_sos < 0 || p->n_sos > APT_MAX_SOS) return fail(ctx, -24, "n_sos=%d outside [0,%d]", p->n_sos, APT_MAX_SOS);
    if (!p->window || !p->freqs) return fail(ctx, -25, "window / freqs tables missing");
    if (p->blk_len < 1 || p->blk_hop < 1 || p->blk_len > 64 || p->blk_hop != p->blk_len)
        return fail(ctx, -26, "unsupported block-energy geometry len=%d hop=%d", p->blk_len, p->blk_hop);
    CUDA_OK(ctx, cudaSetDevice(ctx->device));

    apt_plan* pl = new apt_plan();
    pl->ctx = ctx;
    pl->prm = *p;
    pl->generic = generic;
    pl->td_blocks = td_blocks;
    pl->full_ok = full_ok;
    pl->n_clips = n_clips;
    DevParams& d = pl->dp;
    memset(&d, 0, sizeof(d));
    d.n_fft = p->n_fft; d.hop = p->hop; d.F = F; d.band_lo = p->band_lo; d.K = K; d.M = p->n_modes;
    for (int i = 0; i < APT_MAX_MODES; i++) {
        d.mode_lo[i] = p->mode_lo[i]; d.mode_hi[i] = p->mode_hi[i];
        d.mode_blo[i] = p->mode_band_lo[i]; d.mode_bhi[i] = p->mode_band_hi[i];
        d.mode_w[i] = p->mode_weight[i];
        if (i < p->n_modes && (p->mode_hi[i] >= F || (p->mode_band_hi[i] >= K && p->mode_band_lo[i] <= p->mode_band_hi[i]))) {
            apt_plan_destroy(pl); return fail(ctx, -27, "mode band %d out of range", i);
        }
    }
    d.trk_eta = p->trk_eta; d.trk_alpha = p->trk_scale_alpha; d.trk_1m_alpha = p->trk_one_minus_alpha; d.trk_floor = p->trk_step_floor;
    d.trk_q = p->trk_q; d.trk_nq = p->trk_neg_one_minus_q; d.trk_maxr = p->trk_maxr;
    d.ema_up = p->ema_up; d.ema_down = p->ema_down; d.warm_need = p->warmup_need; d.eps32 = p->eps_f32;
    d.adaptive_q = p->adaptive_q; d.aq_base = p->aq_base; d.aq_min = p->aq_min; d.aq_alpha = p->aq_alpha;
    d.bypass_cls = p->bypass_classifier;
    d.snr_gate = p->snr_gating; d.snr1 = p->snr_gating_snr1;
    d.snr_pow = (p->snr_gating_power > 0.0f && isfinite(p->snr_gating_power)) ? p->snr_gating_power : 1.0f;
    for (int i = 0; i < 4; i++) d.snr_mask[i] = p->snr_mask[i];
    d.pre_smooth = p->pre_smooth_frames > 1 ? p->pre_smooth_frames : 0;
    d.median = p->median_frames > 1 ? p->median_frames : 0;
    if (d.pre_smooth > APT_MAX_PRE_SMOOTH || d.median + ((d.median & 1) == 0 ? 1 : 0) > APT_MAX_MEDIAN) {
        apt_plan_destroy(pl);
        return fail(ctx, -26, "pre_smooth_frames=%d (max %d) / median_frames=%d (max %d) unsupported", p->pre_smooth_frames, APT_MAX_PRE_SMOOTH,
                    p->median_frames, APT_MAX_MEDIAN - 1);
    }
    d.use_norm = p->detector_use_noise_norm; d.ratio_db = p->norm_ratio_db;
    d.bl_q = p->bl_q; d.bl_eta = p->bl_eta; d.bl_alpha = p->bl_scale_alpha; d.bl_floor = p->bl_floor;
    d.norm_enable = p->norm_enable; d.norm_min = p->norm_min_f32;
    d.thr0 = p->thr_primary; d.thr1 = p->thr_m1; d.thr2 = p->thr_m2; d.thr3 = p->thr_m3; d.min_support = p->min_support;
    d.gate_thr = p->td_gate_thr; d.has_ku = p->has_kurt_upper; d.ku = p->kurt_upper;
    d.noise_hi = p->noise_hi; d.mf_noise_max = p->mode_flux_noise_max;
    d.eps64 = p->eps_f64;
    d.blk_len = p->blk_len; d.blk_hop = p->blk_hop; d.blk_pp = p->blk_post_pre; d.blk_smooth = p->blk_smooth;
    d.low_lo = p->low_lo; d.low_hi = p->low_hi; d.rain_lo = p->rain_lo; d.rain_hi = p->rain_hi; d.rolloff = p->rolloff_fraction;
    d.suppressor_bypass = p->suppressor_bypass; d.min_frames = p->clip_rain_min_frames;
    d.gain_mode = p->gain_mode; d.adaptive_gain = p->adaptive_gain; d.gain_freq_smooth = p->gain_freq_smooth;
    d.n_gain_taps = p->n_gain_taps; d.use_lagged = p->use_lagged_noise_psd;
    d.oversub_noise = p->oversub_noise; d.oversub_rain = p->oversub_rain; d.gain_floor = p->gain_floor; d.gain_ceil = p->gain_ceil;
    for (int i = 0; i < APT_MAX_GAIN_TAPS; i++) d.gain_taps[i] = p->gain_taps[i];
    d.alpha_noise = p->alpha_noise; d.om_noise = p->one_minus_alpha_noise; d.alpha_base = p->alpha_base; d.om_base = p->one_minus_alpha_base;
    d.gain_eps = p->gain_eps_f32;
    d.peak_top_p = p->peak_top_p; d.primary_top_m = p->primary_top_m; d.peak_prom_db = p->peak_prominence_db;
    d.peak_min_above_floor = p->peak_min_db_above_floor; d.peak_ratio_min = p->peak_ratio_min;
    d.peak_valid_prom_min = p->peak_valid_prom_min_db; d.peak_valid_prom_max = p->peak_valid_prom_max_db;
    if (p->n_gain_taps < 1 || p->n_gain_taps > APT_MAX_GAIN_TAPS || (p->n_gain_taps & 1) == 0) {
        apt_plan_destroy(pl); return fail(ctx, -32, "n_gain_taps=%d must be odd and <= %d", p->n_gain_taps, APT_MAX_GAIN_TAPS);
    }
    // TD prefilter: n_sos == 0 is run as one identity section
    double sos[APT_MAX_SOS][6];
    int ns = p->n_sos;
    if (ns == 0) {
        ns = 1;
        const double ident[6] = {1, 0, 0, 1, 0, 0};
        memcpy(sos[0], ident, sizeof(ident));
        d.zi[0][0] = d.zi[0][1] = 0.0;
        d.padlen = 0;
    } else {
        for (int s = 0; s < ns; s++) { memcpy(sos[s], p->sos[s], sizeof(double) * 6); d.zi[s][0] = p->zi[s][0]; d.zi[s][1] = p->zi[s][1]; }
        d.padlen = p->padlen;
    }
    d.n_sos = ns;
    for (int s = 0; s < ns; s++) memcpy(d.sos[s], sos[s], sizeof(double) * 6);
    pl->td_ns = ns;

    // offsets
    pl->len.assign(clip_len, clip_len + n_clips);
    pl->samp_off.assign(n_clips + 1, 0); pl->frame_off.assign(n_clips + 1, 0);
    pl->stft_tile_off.assign(n_clips + 1, 0); pl->td_tile_off.assign(n_clips + 1, 0); pl->sel_chunk_off.assign(n_clips + 1, 0); pl->flux_tile_off.assign(n_clips + 1, 0);
    {
        int nl_modes = 0;
        for (int m = 0; m < p->n_modes; m++) nl_modes += std::max(0, p->mode_band_hi[m] - p->mode_band_lo[m] + 1);
        pl->flux_ft = flux_tile_frames(std::max(K, nl_modes));
    }
    for (int c = 0; c < n_clips; c++) {
        const int64_t N = clip_len[c];
        if (N < p->n_fft || N <= d.padlen + 1) { apt_plan_destroy(pl); return fail(ctx, -28, "clip %d too short (%lld samples)", c, (long long)N); }
        const int64_t T = 1 + N / p->hop;
        const int64_t Tloc = 1 + (N - 256) / 128;   // TD tiles are cut in 256 / 128 frames (= 128-sample block pairs) at every geometry
        if (T * (int64_t)std::max(K, 128) >= (int64_t)1 << 31) { apt_plan_destroy(pl); return fail(ctx, -28, "clip %d too long (%lld frames): per-clip plane offsets are 32-bit", c, (long long)T); }
        pl->samp_off[c + 1] = pl->samp_off[c] + N;
        pl->frame_off[c + 1] = pl->frame_off[c] + T;
        pl->stft_tile_off[c + 1] = pl->stft_tile_off[c] + (generic ? (T + stftg_frames_per_cta(p->n_fft) - 1) / stftg_frames_per_cta(p->n_fft) : (T + STFT_TF - 1) / STFT_TF);
        pl->td_tile_off[c + 1] = pl->td_tile_off[c] + std::max<int64_t>(1, (Tloc + TD_FT - 1) / TD_FT);
        pl->sel_chunk_off[c + 1] = pl->sel_chunk_off[c] + (T + DB_CF - 1) / DB_CF;
        pl->flux_tile_off[c + 1] = pl->flux_tile_off[c] + (T + pl->flux_ft - 1) / pl->flux_ft;
    }
    pl->nS = pl->samp_off[n_clips]; pl->nF = pl->frame_off[n_clips];
    if (pl->stft_tile_off[n_clips] > 0x7fffffffLL || pl->td_tile_off[n_clips] > 0x7fffffffLL) { apt_plan_destroy(pl); return fail(ctx, -29, "batch too large for one launch"); }

#define PL_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { int r_ = fail(ctx, -10, "%s failed: %s", #call, cudaGetErrorString(e_)); apt_plan_destroy(pl); return r_; } } while (0)
    PL_OK(upload(pl->d_samp_off, pl->samp_off));
    PL_OK(upload(pl->d_frame_off, pl->frame_off));
    PL_OK(upload(pl->d_stft_tile_off, pl->stft_tile_off));
    PL_OK(upload(pl->d_td_tile_off, pl->td_tile_off));
    PL_OK(upload(pl->d_sel_chunk_off, pl->sel_chunk_off));
    PL_OK(upload(pl->d_flux_tile_off, pl->flux_tile_off));

    if (generic) {
        const int N = p->n_fft;
        std::vector<double> w64(p->window, p->window + N);
        std::vector<float> w32(N);
        for (int i = 0; i < N; i++) w32[i] = (float)w64[i];
        std::vector<cx<double>> tw(N);
        std::vector<cx<float>> twf(N);
        for (int k = 0; k < N; k++) { tw[k] = {cos(2.0 * M_PI * k / N), -sin(2.0 * M_PI * k / N)}; twf[k] = {(float)tw[k].x, (float)tw[k].y}; }
        PL_OK(upload(pl->d_gwin64, w64)); PL_OK(upload(pl->d_gtw64, tw)); PL_OK(upload(pl->d_gwin32, w32)); PL_OK(upload(pl->d_gtw32, twf));
        std::vector<float> fr(p->freqs, p->freqs + F);
        PL_OK(upload(pl->d_freqs, fr));
        if (!full_ok) {
            PL_OK(cudaStreamCreateWithFlags(&pl->s_copy, cudaStreamNonBlocking));
            for (int i = 0; i < apt_plan::N_COMP; i++) PL_OK(cudaStreamCreateWithFlags(&pl->s_comp[i], cudaStreamNonBlocking));
            *out = pl;
            return 0;
        }
    }
    d.td_g = (p->hop % 128 == 0) ? 128 : 64;
    if (td_blocks && full_ok) {
        PL_OK(pl->d_blk_sum.alloc((size_t)(pl->nS / d.td_g + n_clips + 2)));
        PL_OK(pl->d_blk_max.alloc((size_t)(pl->nS / d.td_g + n_clips + 2)));
    }
    // FFT tables
    if (!generic) {
        std::vector<double> w64(p->window, p->window + 256);
        std::vector<float> w32(256);
        for (int i = 0; i < 256; i++) w32[i] = (float)w64[i];
        std::vector<cx<double>> a(128), bq(129);
        std::vector<cx<float>> af(128), bf(129);
        for (int m = 0; m < 128; m++) { a[m] = {cos(2.0 * M_PI * m / 128.0), -sin(2.0 * M_PI * m / 128.0)}; af[m] = {(float)a[m].x, (float)a[m].y}; }
        for (int k = 0; k <= 128; k++) { bq[k] = {cos(2.0 * M_PI * k / 256.0), -sin(2.0 * M_PI * k / 256.0)}; bf[k] = {(float)bq[k].x, (float)bq[k].y}; }
        PL_OK(upload(pl->d_win64, w64)); PL_OK(upload(pl->d_tw128_64, a)); PL_OK(upload(pl->d_tw256_64, bq));
        PL_OK(upload(pl->d_win32, w32)); PL_OK(upload(pl->d_tw128_32, af)); PL_OK(upload(pl->d_tw256_32, bf));
        std::vector<float> fr(p->freqs, p->freqs + F);
        PL_OK(upload(pl->d_freqs, fr));
    }
    if (p->fft_f64 == 2 && !generic) {
        // B[n][col] = window[n] * (cos, -sin)(2 pi k n / 256) for the band bins k, times 1024, as two fp16 limbs, laid out
        // as the K-major SWIZZLE_128B shared-memory image the tensor-core kernel copies in: [limb][chunk of 64 n][row col]
        if (2 * K > TC_N) { apt_plan_destroy(pl); return fail(ctx, -22, "tensor-core DFT: at most %d band bins", TC_N / 2); }
        std::vector<unsigned char> img(TC_B_BYTES, 0);
        for (int col = 0; col < 2 * K; col++) {
            const int k = p->band_lo + (col >> 1);
            for (int n = 0; n < 256; n++) {
                const double ang = 2.0 * M_PI * (double)((k * n) & 255) / 256.0;
                const double v = p->window[n] * ((col & 1) ? -sin(ang) : cos(ang)) * (double)TC_BSCALE;
                const __half h1 = __float2half_rn((float)v);
                const __half h2 = __float2half_rn((float)(v - (double)__half2float(h1)));
                const int chunk = n / TC_KC, kk = n % TC_KC;
                const int off = tc_swz(col, kk >> 3) + (kk & 7) * 2;
                memcpy(&img[(size_t)(0 * TC_NCHUNK + chunk) * TC_B_TILE + off], &h1, 2);
                memcpy(&img[(size_t)(1 * TC_NCHUNK + chunk) * TC_B_TILE + off], &h2, 2);
            }
        }
        PL_OK(upload(pl->d_tc_B, img));
        PL_OK(pl->d_tc_err.alloc(1));
        PL_OK(cudaMemset(pl->d_tc_err.p, 0, sizeof(int)));
    }
    // TD tables
    {
        const int halo = (p->blk_post_pre + 2) * p->blk_hop + p->blk_len;
        const int lb = TD_FT * 128 + 256 + 2 * halo + 2 * TD_WARM + 128;
        if (halo > 128 || lb > TD_LB) { apt_plan_destroy(pl); return fail(ctx, -26, "block-energy geometry needs a %d-sample tile buffer (max %d)", lb, TD_LB); }
        std::vector<double> Apow, H;
        build_td_tables(*p, ns, sos, TD_CHUNK, Apow, H);
        PL_OK(upload(pl->d_Apow, Apow)); PL_OK(upload(pl->d_H, H));
        pl->tdt.Alin = pl->d_Apow.p; pl->tdt.H = pl->d_H.p; pl->tdt.chunk = TD_CHUNK; pl->tdt.lb_max = TD_LB; pl->tdt.halo = halo;
        {   // the tile warm-up (TD_WARM samples) and the one-level warp scan both rely on the filter's memory
            // being numerically gone after TD_WARM samples: check A^(TD_WARM / TD_CHUNK)
            const int dim = 2 * ns, e = TD_WARM / TD_CHUNK;
            double mx = 0.0;
            for (int i = 0; i < dim * dim; i++) mx = std::max(mx, fabs(Apow[(size_t)i * 32 + e]));
            if (mx > 1e-15) { apt_plan_destroy(pl); return fail(ctx, -26, "TD prefilter decays too slowly for the %d-sample tile warm-up (|A^%d| = %.3g)", TD_WARM, e, mx); }
        }
        memset(pl->tdt.Hc, 0, sizeof(pl->tdt.Hc)); memset(pl->tdt.Adc, 0, sizeof(pl->tdt.Adc));
        if (ns <= 2) {
            const int dim = 2 * ns;
            for (int m = 0; m < TD_CHUNK; m++) for (int r = 0; r < dim; r++) pl->tdt.Hc[m * dim + r] = H[(size_t)m * dim + r];
            for (int k = 0; k < 5; k++) for (int i = 0; i < dim * dim; i++) pl->tdt.Adc[k * 16 + i] = Apow[(size_t)i * 32 + (1 << k)];
        }
        pl->tdt.env_cap = (TD_FT * 128 + 256 + 2 * halo) / std::max(1, p->blk_hop) + 8;
        if (ns == 2) {   // tables of the float32 gate kernel (chunk = 32 samples)
            std::vector<double> Ap32, H32;
            build_td_tables(*p, ns, sos, TDF_CH, Ap32, H32);
            std::vector<float> tab(TDF_TAB_N, 0.0f);
            for (int m = 0; m < TDF_CH; m++) for (int r = 0; r < 4; r++) tab[TDF_TAB_H + m * 4 + r] = (float)H32[(size_t)m * 4 + r];
            for (int k = 0; k < 5; k++) for (int i = 0; i < 16; i++) tab[TDF_TAB_A2K + k * 16 + i] = (float)Ap32[(size_t)i * 32 + (1 << k)];
            for (int e = 0; e < 32; e++) for (int i = 0; i < 16; i++) tab[TDF_TAB_APOS + e * 20 + i] = (float)Ap32[(size_t)i * 32 + e];
            double mx = 0.0;     // what is left of a state after the warm-up (12 chunks) must be far below float32 resolution
            for (int i = 0; i < 16; i++) mx = std::max(mx, fabs(Ap32[(size_t)i * 32 + TDF_WARM / TDF_CH]));
            PL_OK(upload(pl->d_tdf_tab, tab));
            memset(&pl->tdf, 0, sizeof(pl->tdf));
            for (int s_ = 0; s_ < 2; s_++) for (int j = 0; j < 6; j++) pl->tdf.c[s_][j] = (float)sos[s_][j];
            pl->tdf.tab = pl->d_tdf_tab.p;
            pl->tdf.eps = (float)p->eps_f64;
            pl->tdf.thr = p->td_gate_thr;
            pl->tdf_ok = mx < 1e-9;
        }
        pl->td_smem = td_smem_bytes(ns, pl->tdt.env_cap);
        pl->td_smem_f32 = td_smem_bytes(ns, pl->tdt.env_cap, sizeof(float));
        for (int i = 0; i < TD_CHUNK * 4; i++) pl->tdt.Hcf[i] = (float)pl->tdt.Hc[i];
        for (int i = 0; i < 5 * 16; i++) pl->tdt.Adcf[i] = (float)pl->tdt.Adc[i];
    }
    // scratch
    PL_OK(pl->d_Pband.alloc((size_t)pl->nF * K));
    PL_OK(pl->d_n2.alloc((size_t)pl->nF * K));
    PL_OK(pl->d_td.alloc((size_t)pl->nF * APT_N_TD_FEATURES));
    PL_OK(pl->d_dbsum.alloc((size_t)pl->sel_chunk_off[n_clips]));
    pl->mf_stride = (p->n_modes + 1 <= 8) ? 8 : 16;
    PL_OK(pl->d_mf.alloc((size_t)pl->nF * pl->mf_stride));
    {   // pass-1 lane tables
        Trk1Tab& tm = pl->tab_modes; Trk1Tab& ta = pl->tab_all;
        memset(&tm, 0, sizeof(tm)); memset(&ta, 0, sizeof(ta));
        int nl = 0;
        std::vector<unsigned short> lm, la(K);
        for (int m = 0; m < p->n_modes; m++) {
            const int lo = p->mode_band_lo[m], hi = p->mode_band_hi[m];
            const int n = hi >= lo ? hi - lo + 1 : 0;
            tm.mode_l0[m] = nl; tm.mode_n[m] = n;
            ta.mode_l0[m] = n ? lo : 0; ta.mode_n[m] = n;
            for (int i = 0; i < n; i++) {
                if (nl >= SEQ_KMAX && !generic) { apt_plan_destroy(pl); return fail(ctx, -27, "mode bands cover more than %d bins", SEQ_KMAX); }
                if (nl < SEQ_KMAX) tm.lane_bin[nl] = (unsigned char)(lo + i);
                lm.push_back((unsigned short)(lo + i));
                nl++;
            }
        }
        tm.n_lanes = nl; tm.nls = std::max(8, (nl + 7) / 8 * 8);
        ta.n_lanes = K;  ta.nls = (K + 7) / 8 * 8;
        PL_OK(pl->d_nl.alloc((size_t)pl->nF * tm.nls));   // the all-bins plane (debug outputs) is allocated on first use
        for (int k = 0; k < K; k++) { if (k < SEQ_KMAX) ta.lane_bin[k] = (unsigned char)k; la[k] = (unsigned short)k; }
        // frame sizes whose tables outgrow the kernel-parameter arrays (255 as a bin index, SEQ_KMAX lanes) read them from global memory
        if (nl > SEQ_KMAX || K > SEQ_KMAX) {
            PL_OK(upload(pl->d_lane_modes, lm)); PL_OK(upload(pl->d_lane_all, la));
            tm.lane_bin_g = pl->d_lane_modes.p; ta.lane_bin_g = pl->d_lane_all.p;
        }
    }
    PL_OK(pl->d_sel.alloc(n_clips));
    PL_OK(pl->d_hist.alloc((size_t)n_clips * 2 * SEL_BINS));
    PL_OK(pl->d_gate.alloc((size_t)pl->nF));
    {   // flagged-tile lists: one slice of (tiles per segment) x clips per time segment; segments x tiles per segment is
        // below twice the tiles of the longest clip
        int64_t mx = 1;
        for (int c = 0; c < n_clips; c++) mx = std::max(mx, pl->td_tile_off[c + 1] - pl->td_tile_off[c]);
        pl->td_list_cap = 2 * mx * n_clips + 64;
    }
    PL_OK(pl->d_td_list.alloc((size_t)pl->td_list_cap));
    PL_OK(pl->d_td_cnt.alloc(apt_plan::MAX_SEG));
    if (const char* e = getenv("APT_TD_FAST")) pl->td_fast = atoi(e);
    if (const char* e = getenv("APT_TD_GUARD")) pl->td_guard = (float)atof(e);
    pl->st_stride = std::max(K, pl->tab_modes.n_lanes);
    PL_OK(pl->d_st_trk1.alloc((size_t)n_clips * pl->st_stride));
    PL_OK(pl->d_st_trk2.alloc((size_t)n_clips * pl->st_stride));
    PL_OK(pl->d_st_base.alloc((size_t)n_clips * (APT_MAX_MODES + 1)));
    PL_OK(pl->d_st_aq.alloc((size_t)n_clips));
    if (d.pre_smooth) {
        PL_OK(pl->d_Psm.alloc((size_t)pl->nF * K));
        PL_OK(pl->d_st_ps.alloc((size_t)n_clips * pl->st_stride * PS_STATE));
    }
    if (d.pre_smooth || d.median) PL_OK(pl->d_N1raw.alloc((size_t)pl->nF * K));
    if (d.median) { PL_OK(pl->d_N1med.alloc((size_t)pl->nF * K)); PL_OK(pl->d_N2raw.alloc((size_t)pl->nF * K)); }
    {   // candidate lists of the median select: an eighth of the plane per clip (a 1/16 dB bin holds ~1 % of a clip's
        // values; clips that overflow fall back to the full radix select)
        pl->cand_off.assign(n_clips + 1, 0);
        for (int c = 0; c < n_clips; c++) {
            const int64_t n = (pl->frame_off[c + 1] - pl->frame_off[c]) * (int64_t)K;
            pl->cand_off[c + 1] = pl->cand_off[c] + std::max<int64_t>(1024, (n / 8 + 31) / 32 * 32);
        }
        PL_OK(upload(pl->d_cand_off, pl->cand_off));
        PL_OK(pl->d_cand.alloc((size_t)pl->cand_off[n_clips]));
    }
    PL_OK(pl->d_counter.alloc(64));
    pl->scratch_bytes = sizeof(float) * ((size_t)pl->nF * K * 2 + (size_t)pl->nF * APT_N_TD_FEATURES + (size_t)pl->nF * pl->mf_stride +
                                         (size_t)pl->nF * pl->tab_modes.nls + (size_t)pl->cand_off[n_clips] + (size_t)n_clips * K * 8) +
                        sizeof(double) * (size_t)pl->sel_chunk_off[n_clips] +
                        (size_t)n_clips * (sizeof(SelState) + 2 * SEL_BINS * sizeof(uint32_t));
    PL_OK(cudaStreamCreateWithFlags(&pl->s_copy, cudaStreamNonBlocking));
    for (int i = 0; i < apt_plan::N_COMP; i++) PL_OK(cudaStreamCreateWithFlags(&pl->s_comp[i], cudaStreamNonBlocking));
    {
        int lo = 0, hi = 0;
        PL_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // lo = least, hi = greatest priority
        for (int k = 0; k < apt_plan::SK_N; k++) {
            const int prio = (k == apt_plan::SK_STFT || k == apt_plan::SK_TD) ? lo : hi;
            PL_OK(cudaStreamCreateWithPriority(&pl->s_kind[k], cudaStreamNonBlocking, prio));
            for (int i = 0; i < apt_plan::MAX_SEG; i++) PL_OK(cudaEventCreateWithFlags(&pl->ev_seg[k][i], cudaEventDisableTiming));
        }
        PL_OK(cudaEventCreateWithFlags(&pl->ev_fork, cudaEventDisableTiming));
        if (const char* e = getenv("APT_PIPELINE")) pl->pipeline = atoi(e);
    }
#undef PL_OK
    *out = pl;
    return 0;
}

void apt_plan_destroy(apt_plan_t* plan) {
    if (!plan) return;
    if (plan->s_copy) cudaStreamDestroy(plan->s_copy);
    for (int i = 0; i < apt_plan::N_COMP; i++) if (plan->s_comp[i]) cudaStreamDestroy(plan->s_comp[i]);
    for (int k = 0; k < apt_plan::SK_N; k++) {
        if (plan->s_kind[k]) cudaStreamDestroy(plan->s_kind[k]);
        for (int i = 0; i < apt_plan::MAX_SEG; i++) if (plan->ev_seg[k][i]) cudaEventDestroy(plan->ev_seg[k][i]);
    }
    if (plan->ev_fork) cudaEventDestroy(plan->ev_fork);
    if (plan->tr_origin) cudaEventDestroy(plan->tr_origin);
    if (plan->tr_end) cudaEventDestroy(plan->tr_end);
    for (int k = 0; k < apt_plan::SK_N; k++) for (int i = 0; i < apt_plan::MAX_SEG; i++) for (int j = 0; j < 2; j++)
        if (plan->tr_ev[k][i][j]) cudaEventDestroy(plan->tr_ev[k][i][j]);
    for (int i = 0; i < apt_plan::N_RING; i++) if (plan->ring[i]) cudaFreeHost(plan->ring[i]);
    for (auto& m : plan->marks) cudaEventDestroy(m.second);
    delete plan;
}

int apt_plan_offsets(const apt_plan_t* plan, int64_t* so, int64_t* fo) {
    if (!plan) return -1;
    if (so) memcpy(so, plan->samp_off.data(), sizeof(int64_t) * (plan->n_clips + 1));
    if (fo) memcpy(fo, plan->frame_off.data(), sizeof(int64_t) * (plan->n_clips + 1));
    return 0;
}
int64_t apt_plan_total_frames(const apt_plan_t* plan) { return plan ? plan->nF : -1; }
int64_t apt_plan_total_samples(const apt_plan_t* plan) { return plan ? plan->nS : -1; }
int64_t apt_plan_scratch_bytes(const apt_plan_t* plan) { return plan ? (int64_t)plan->scratch_bytes : -1; }
int apt_plan_last_launches(const apt_plan_t* plan) { return plan ? plan->last_launches : -1; }

int apt_plan_enable_timing(apt_plan_t* plan, int enable) {
    if (!plan) return -1;
    plan->timing = enable != 0;
    for (auto& m : plan->marks) cudaEventDestroy(m.second);
    plan->marks.clear();
    for (int i = 0; i < APT_N_KERNELS; i++) plan->kernel_ms[i] = 0.0f;
    return 0;
}

int apt_plan_tc_error(apt_plan_t* plan) {
    if (!plan) return -1;
    if (!plan->d_tc_err.p) return 0;
    int h = 0;
    if (cudaMemcpy(&h, plan->d_tc_err.p, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -10;
    return h;
}

int apt_plan_enable_trace(apt_plan_t* plan, int enable) {
    if (!plan) return -1;
    if (enable && !plan->tr_origin) {
        if (cudaEventCreate(&plan->tr_origin) != cudaSuccess || cudaEventCreate(&plan->tr_end) != cudaSuccess) return -10;
        for (int k = 0; k < apt_plan::SK_N; k++) for (int i = 0; i < apt_plan::MAX_SEG; i++) for (int j = 0; j < 2; j++)
            if (cudaEventCreate(&plan->tr_ev[k][i][j]) != cudaSuccess) return -10;
    }
    plan->trace = enable != 0;
    plan->tr_nseg = 0;
    return 0;
}

int apt_plan_trace(apt_plan_t* plan, float* out_ms, int* n_seg, float* total_ms) {
    if (!plan || !out_ms || !n_seg || !total_ms) return -1;
    *n_seg = plan->tr_nseg;
    *total_ms = 0.0f;
    if (!plan->tr_nseg) return 0;
    cudaEventElapsedTime(total_ms, plan->tr_origin, plan->tr_end);
    for (int k = 0; k < apt_plan::SK_N; k++) for (int i = 0; i < plan->tr_nseg; i++) for (int j = 0; j < 2; j++) {
        float ms = -1.0f;
        if (cudaEventElapsedTime(&ms, plan->tr_origin, plan->tr_ev[k][i][j]) != cudaSuccess) { ms = -1.0f; cudaGetLastError(); }
        out_ms[(k * apt_plan::MAX_SEG + i) * 2 + j] = ms;
    }
    return 0;
}

int apt_plan_kernel_ms(apt_plan_t* plan, float* out_ms) {
    if (!plan || !out_ms) return -1;
    // caller has synchronised the stream; accumulate the intervals between consecutive marks
    for (size_t i = 0; i + 1 < plan->marks.size(); i++) {
        const int id = plan->marks[i].first;
        if (id < 0) continue;
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, plan->marks[i].second, plan->marks[i + 1].second) == cudaSuccess) plan->kernel_ms[id] += ms;
    }
    for (auto& m : plan->marks) cudaEventDestroy(m.second);
    plan->marks.clear();
    for (int i = 0; i < APT_N_KERNELS; i++) { out_ms[i] = plan->kernel_ms[i]; plan->kernel_ms[i] = 0.0f; }
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------
// grid of a tiled kernel over the tiles [tile0, tile0 + per_seg) of clips [clip0, clip0 + n): x = tiles of the
// longest clip that fall into the range, y = clips.  x == 0: nothing to launch.
static dim3 seg_grid(const std::vector<int64_t>& off, int clip0, int n, int64_t tile0, int64_t per_seg) {
    int64_t mx = 0;
    for (int c = clip0; c < clip0 + n; c++) mx = std::max(mx, off[c + 1] - off[c]);
    const int64_t x = std::max<int64_t>(0, std::min(per_seg, mx - tile0));
    return dim3((unsigned)x, (unsigned)n, 1);
}

// Tiled kernels (STFT, TD) can walk several tiles of a clip per CTA (grid.x < tiles), which pays the per-CTA set-up once.
// Measured (profiles/r2): with 8 waves of long-lived CTAs the kernels lose more to lock-step phases (every resident
// CTA staging, then every CTA in the FFT) than the set-up costs -- issue utilisation 64 % -> 57 % (STFT), 80 % -> 65 %
// (TD) -- so the default is one tile per CTA; APT_PERSIST_WAVES=n selects n waves of persistent CTAs.
static dim3 persistent_grid(const apt_plan* pl, dim3 tiles, int ctas_per_sm) {
    int waves = 0;
    if (const char* e = getenv("APT_PERSIST_WAVES")) waves = atoi(e);
    if (tiles.x == 0 || waves <= 0) return tiles;
    const int64_t want = (int64_t)waves * ctas_per_sm * pl->ctx->sm_count;
    const int64_t gx = std::max<int64_t>(1, std::min<int64_t>(tiles.x, (want + tiles.y - 1) / tiles.y));
    return dim3((unsigned)gx, tiles.y, 1);
}

template <typename T, typename PCM>
static cudaError_t launch_stft(apt_plan* pl, const Batch& b, dim3 grid, const PCM* pcm, const StftOut& so, cudaStream_t st) {
    if (grid.x == 0) return cudaSuccess;
    FftTables<T> tab;
    if constexpr (sizeof(T) == 8) { tab.win = pl->d_win64.p; tab.tw128 = pl->d_tw128_64.p; tab.tw256 = pl->d_tw256_64.p; }
    else { tab.win = pl->d_win32.p; tab.tw128 = pl->d_tw128_32.p; tab.tw256 = pl->d_tw256_32.p; }
    const size_t smem = stft_smem_bytes<T>();
    auto kern = stft256_kernel<T, PCM>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, STFT_NT, smem, st>>>(pl->dp, b, pcm, pl->d_stft_tile_off.p, tab, so);
    pl->last_launches++;
    return cudaGetLastError();
}

// tensor-core DFT (fft mode 2, int16 input, band plane / band energies only)
static cudaError_t launch_tcdft(apt_plan* pl, const Batch& b, const int16_t* pcm, const StftOut& so, cudaStream_t st) {
    TcParams q;
    memset(&q, 0, sizeof(q));
    q.Bimg = pl->d_tc_B.p; q.P_band = so.P_band; q.band_energy = so.band_energy; q.nF = pl->nF;
    q.K = pl->dp.K; q.band_lo = pl->dp.band_lo; q.n_modes = pl->dp.M;
    for (int m = 0; m < APT_MAX_MODES; m++) { q.mode_blo[m] = pl->dp.mode_blo[m]; q.mode_bhi[m] = pl->dp.mode_bhi[m]; }
    q.eps = (float)pl->dp.eps64; q.error_flag = pl->d_tc_err.p;
    int64_t maxT = 1;
    for (int c = b.clip0; c < b.clip0 + b.n_clips; c++) maxT = std::max(maxT, pl->frame_off[c + 1] - pl->frame_off[c]);
    const int64_t t_hi = std::min<int64_t>(maxT, b.tb);
    const int64_t tiles = std::max<int64_t>(0, (t_hi + TC_M - 1) / TC_M - b.ta / TC_M);
    if (tiles == 0) return cudaSuccess;
    // one CTA per SM (214 KB of shared memory): as many CTAs per clip as fit in ONE wave over the launch's clips
    const int64_t gx = std::max<int64_t>(1, std::min<int64_t>(tiles, pl->ctx->sm_count / b.n_clips));
    auto kern = tcdft256_kernel<int16_t>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    if (e != cudaSuccess) return e;
    kern<<<dim3((unsigned)gx, (unsigned)b.n_clips), TC_NT, TC_SMEM, st>>>(b, pcm, pl->d_stft_tile_off.p, q);
    pl->last_launches++;
    return cudaGetLastError();
}
// STFT of the 256-sample geometry in the plan's arithmetic: 1 = float64 FFT (the reference's), 0 = float32 FFT,
// 2 = tensor-core DFT (tolerance path; int16 input, band plane / band energies only -- anything else runs the float32 FFT)
template <typename PCM>
static cudaError_t launch_stft_mode(apt_plan* pl, const Batch& b, dim3 grid, const PCM* pcm, const StftOut& so, cudaStream_t st) {
    const int mode = pl->prm.fft_f64;
    if (mode == 2) {
        if constexpr (sizeof(PCM) == 2) {
            if (!so.S && !so.P && !so.raw && pl->d_tc_B.p) return launch_tcdft(pl, b, pcm, so, st);
        }
        return launch_stft<float, PCM>(pl, b, grid, pcm, so, st);
    }
    return mode ? launch_stft<double, PCM>(pl, b, grid, pcm, so, st) : launch_stft<float, PCM>(pl, b, grid, pcm, so, st);
}

template <typename T, typename PCM>
static cudaError_t launch_stft_generic(apt_plan* pl, const Batch& b, const PCM* pcm, const StftOut& so, cudaStream_t st) {
    FftTablesG<T> tab;
    if constexpr (sizeof(T) == 8) { tab.win = pl->d_gwin64.p; tab.tw = pl->d_gtw64.p; }
    else { tab.win = pl->d_gwin32.p; tab.tw = pl->d_gtw32.p; }
    const size_t smem = stftg_smem_bytes<T>(pl->dp.n_fft);
    auto kern = stft_generic_kernel<T, PCM>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<seg_grid(pl->stft_tile_off, b.clip0, b.n_clips, 0, INT64_MAX), STFTG_NT, smem, st>>>(pl->dp, b, pcm, tab, so, stftg_frames_per_cta(pl->dp.n_fft));
    pl->last_launches++;
    return cudaGetLastError();
}

template <int NS, typename PCM, typename R>
static cudaError_t launch_td_ns(apt_plan* pl, const Batch& b, dim3 grid, const PCM* pcm, const TdOut& to, cudaStream_t st) {
    if (grid.x == 0) return cudaSuccess;
    auto kern = td_features_kernel<NS, PCM, R>;
    const size_t smem = sizeof(R) == 8 ? pl->td_smem : pl->td_smem_f32;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, TD_NT, smem, st>>>(pl->dp, b, pcm, pl->d_td_tile_off.p, pl->tdt, to);
    pl->last_launches++;
    return cudaGetLastError();
}
template <typename PCM, typename R>
static cudaError_t launch_td(apt_plan* pl, const Batch& b, dim3 grid, const PCM* pcm, const TdOut& to, cudaStream_t st) {
    switch (pl->td_ns) {
        case 1: return launch_td_ns<1, PCM, R>(pl, b, grid, pcm, to, st);
        case 2: return launch_td_ns<2, PCM, R>(pl, b, grid, pcm, to, st);
        case 3: return launch_td_ns<3, PCM, R>(pl, b, grid, pcm, to, st);
        case 4: return launch_td_ns<4, PCM, R>(pl, b, grid, pcm, to, st);
    }
    return cudaErrorInvalidValue;
}
template <int NS, typename PCM>
static cudaError_t launch_td_recheck_ns(apt_plan* pl, const Batch& b, const PCM* pcm, const TdOut& to, cudaStream_t st) {
    auto kern = td_recheck_kernel<NS, PCM>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->td_smem);
    if (e != cudaSuccess) return e;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(to.list_cap, 2 * (int64_t)pl->ctx->sm_count));
    kern<<<grid, TD_NT, pl->td_smem, st>>>(pl->dp, b, pcm, pl->d_td_tile_off.p, pl->tdt, to);
    pl->last_launches++;
    return cudaGetLastError();
}
template <typename PCM>
static cudaError_t launch_td_recheck(apt_plan* pl, const Batch& b, const PCM* pcm, const TdOut& to, cudaStream_t st) {
    switch (pl->td_ns) {
        case 1: return launch_td_recheck_ns<1, PCM>(pl, b, pcm, to, st);
        case 2: return launch_td_recheck_ns<2, PCM>(pl, b, pcm, to, st);
        case 3: return launch_td_recheck_ns<3, PCM>(pl, b, pcm, to, st);
        case 4: return launch_td_recheck_ns<4, PCM>(pl, b, pcm, to, st);
    }
    return cudaErrorInvalidValue;
}

// Frames per time segment of the pipelined run: a common multiple of every tile size on the path (STFT 32, TD 56,
// flux 256, dB chunk 896 frames -> 1792), sized for about `want` segments over the longest clip.
constexpr int SEG_UNIT = 1792;
static_assert(SEG_UNIT % STFT_TF == 0 && SEG_UNIT % TD_FT == 0 && SEG_UNIT % FLUX_FT == 0 && SEG_UNIT % DB_CF == 0, "segment unit");

template <typename PCM>
static int run_range(apt_plan* pl, int stages, int clip0, int n_clips, const PCM* pcm, const apt_out_t* out, cudaStream_t st,
                     int want_seg = 16) {
    apt_ctx* ctx = pl->ctx;
    const DevParams& d = pl->dp;
    const bool full = (stages & APT_STAGE_FULL) != 0;
    // ---- every argument check and lazy allocation happens before the first launch
    if (full && (!out->frame_class || !out->rain_conf || !out->noise_conf || !out->event_idx || !out->event_count || !out->clip_stats))
        return fail(ctx, -30, "full pipeline requires frame_class, rain_conf, noise_conf, event_idx, event_count, clip_stats buffers");
    if (!full && !(out->band_energy || out->P || out->S || out->raw))
        return fail(ctx, -31, "features stage requires at least one of band_energy / P / S / raw");
    if (full && !pl->full_ok)
        return fail(ctx, -34, "the full pipeline needs n_fft=256 / hop=128, or a hop that is a multiple of 64 (and no kurtosis gate); "
                              "n_fft=%d hop=%d supports the features stage", d.n_fft, d.hop);
    if (full && pl->td_blocks && (out->G || out->S_hat || out->y || out->peak_ratio || out->peak_gate_score || out->peak_valid_count ||
                                out->peak_count_by_mode || out->ratio_med))
        return fail(ctx, -34, "n_fft=%d hop=%d: the gain / resynthesis planes and the peak features exist at n_fft=256 / hop=128 only",
                    d.n_fft, d.hop);
    const bool want_peaks = out->peak_ratio || out->peak_gate_score || out->peak_valid_count || out->peak_count_by_mode;
    const bool want_gain = out->G || out->S_hat || out->y;
    if (full) {
        if (want_peaks && !(out->peak_ratio && out->peak_gate_score && out->peak_valid_count && out->peak_count_by_mode))
            return fail(ctx, -35, "the four peak-feature buffers must be given together");
        if (want_gain && !d.suppressor_bypass) {
            if (!out->G) return fail(ctx, -33, "S_hat / y require the G buffer");
            if (out->y && !out->S_hat) return fail(ctx, -33, "y requires the S_hat buffer");
            if (out->S_hat && !out->S) return fail(ctx, -33, "S_hat requires the S buffer");
        }
    }
    float* D_plane = out->D;
    if (full && want_peaks && !D_plane) {   // the peak features read the detector input of every band bin
        if (!pl->d_Dscr.p) CUDA_OK(ctx, pl->d_Dscr.alloc((size_t)pl->nF * d.K));
        D_plane = pl->d_Dscr.p;
    }
    const bool dbg = full && (out->det_noise_psd || out->det_noise_lag || D_plane);
    if (dbg && !pl->d_nl_all.p) CUDA_OK(ctx, pl->d_nl_all.alloc((size_t)pl->nF * pl->tab_all.nls));

    Batch b{clip0, n_clips, pl->d_samp_off.p, pl->d_frame_off.p, 0, 0, INT32_MAX, INT32_MAX};
    StftOut so;
    so.S = out->S; so.P = out->P; so.P_band = full ? pl->d_Pband.p : nullptr; so.band_energy = out->band_energy;
    so.raw = out->raw; so.freqs = pl->d_freqs.p; so.nF = pl->nF;
    if (pl->generic && !full) {
        pl->mark(APT_KERNEL_STFT, st);
        cudaError_t eg = pl->prm.fft_f64 ? launch_stft_generic<double, PCM>(pl, b, pcm, so, st) : launch_stft_generic<float, PCM>(pl, b, pcm, so, st);
        if (eg != cudaSuccess) return fail(ctx, -11, "stft launch failed: %s", cudaGetErrorString(eg));
        pl->mark(-1, st);
        return 0;
    }
    if (!full) {
        pl->mark(APT_KERNEL_STFT, st);
        const dim3 g = persistent_grid(pl, seg_grid(pl->stft_tile_off, clip0, n_clips, 0, INT64_MAX), 3);
        cudaError_t e = launch_stft_mode<PCM>(pl, b, g, pcm, so, st);
        if (e != cudaSuccess) return fail(ctx, -11, "stft launch failed: %s", cudaGetErrorString(e));
        pl->mark(-1, st);
        return 0;
    }

    // ---- the full pipeline.  The time axis of every clip is cut into segments; all kernels of one kind run on their
    // own stream in segment order (the serial kernels carry their state from segment to segment), and events chain
    // the kinds inside a segment:  stft -> trk1 -> flux -> base -> decide (<- td) -> trk2 -> dbsum.
    // So the latency-bound serial kernels of segment s overlap each other's neighbours and the wide kernels of the
    // segments behind them, instead of each paying its whole chain (frames x dependent-issue latency) alone.
    // With per-kernel timing on (or APT_PIPELINE=0) everything runs on the caller's stream in one segment.
    int64_t maxT = 1;
    for (int c = clip0; c < clip0 + n_clips; c++) maxT = std::max(maxT, pl->frame_off[c + 1] - pl->frame_off[c]);
    // (the generic frame sizes run in one segment on the caller's stream: their TD tiles are not aligned with frame ranges)
    const bool piped = want_seg > 1 && pl->pipeline && !pl->timing && !pl->td_blocks;
    int seg_frames, n_seg;
    if (piped) {
        int want = want_seg;
        if (const char* e = getenv("APT_SEGMENTS")) want = std::max(1, std::min((int)apt_plan::MAX_SEG, atoi(e)));
        const int64_t units = (maxT + SEG_UNIT - 1) / SEG_UNIT;
        const int64_t upseg = std::max<int64_t>(1, (units + want - 1) / want);
        seg_frames = (int)std::min<int64_t>(upseg * SEG_UNIT, INT32_MAX / 2);
        n_seg = (int)((maxT + seg_frames - 1) / seg_frames);
        if (n_seg > apt_plan::MAX_SEG) return fail(ctx, -29, "internal: %d time segments", n_seg);
    } else {
        seg_frames = (int)std::min<int64_t>((maxT + SEG_UNIT - 1) / SEG_UNIT * SEG_UNIT, INT32_MAX / 2);
        n_seg = 1;
    }
    cudaStream_t S[apt_plan::SK_N];
    for (int k = 0; k < apt_plan::SK_N; k++) S[k] = piped ? pl->s_kind[k] : st;
    // The two bulk kernels share ONE stream (STFT of a segment, then its TD kernel): run side by side they take 95 ms
    // for the 1 000-clip batch, back to back 67 ms (both live on the FP64 pipe and on most of an SM's shared memory;
    // profiles/r2/trace_*.txt) -- the chain kernels are what overlaps them.
    // Schedule.  Large batches (the serial kernels alone hold >= 12 warps per SM: they are throughput-bound themselves,
    // and every CTA of theirs that sits on an SM takes a third of its registers away from the bulk kernels) run the
    // bulk kernels of all segments first and pipeline only the chain behind them ("two-phase": 89.7 against 92.2 ms at
    // 1 000 clips); smaller batches pipeline everything (125 clips: 14.5 against 17.4 ms).  profiles/r2/variants_*.txt
    int td_own = 0;
    int two_phase = ((double)n_clips * d.K / 32.0 / std::max(1, ctx->sm_count)) >= 12.0 ? 1 : 0;
    if (const char* e = getenv("APT_TD_OWN_STREAM")) td_own = atoi(e);
    if (const char* e = getenv("APT_TWO_PHASE")) two_phase = atoi(e);
    if (!td_own) S[apt_plan::SK_TD] = S[apt_plan::SK_STFT];
    // record on the producing kind's stream / wait on the consuming kind's stream (no-ops on one stream)
    auto rec = [&](int kind, int sg) -> cudaError_t {
        if (!piped) return cudaSuccess;
        if (pl->trace) cudaEventRecord(pl->tr_ev[kind][sg][1], S[kind]);
        return cudaEventRecord(pl->ev_seg[kind][sg], S[kind]);
    };
    auto wait = [&](int kind, int on, int sg) -> cudaError_t { return piped ? cudaStreamWaitEvent(S[kind], pl->ev_seg[on][sg], 0) : cudaSuccess; };

    const Trk1Tab& tab = dbg ? pl->tab_all : pl->tab_modes;
    float* nl_plane = dbg ? pl->d_nl_all.p : pl->d_nl.p;
    float* n2_plane = out->noise_psd ? out->noise_psd : pl->d_n2.p;
    // optional smoothing: the trackers read the smoothed power; pass 1's result goes through the median and a lag / clamp
    // pass of its own, pass 2's through the median
    const bool PS = d.pre_smooth > 1, MED = d.median > 1, post1 = PS || MED;
    const float* trk_in = PS ? pl->d_Psm.p : pl->d_Pband.p;
    float* n1_raw = !post1 ? out->det_noise_psd : (MED ? pl->d_N1raw.p : (out->det_noise_psd ? out->det_noise_psd : pl->d_N1raw.p));
    float* n1_fin = MED ? (out->det_noise_psd ? out->det_noise_psd : pl->d_N1med.p) : n1_raw;
    float* n2_raw = MED ? pl->d_N2raw.p : n2_plane;
    auto frame_lane_grid = [&](const Batch& bb, int lanes_) {
        const int64_t fr = std::max<int64_t>(0, std::min<int64_t>(maxT - bb.ta, n_seg == 1 ? maxT : (int64_t)seg_frames));
        return dim3((unsigned)((fr * lanes_ + 255) / 256), (unsigned)n_clips);
    };
    uint32_t* hist = pl->d_hist.p;
    const size_t hist_bytes = sizeof(uint32_t) * (size_t)n_clips * 2 * SEL_BINS;
    TdOut to;
    memset(&to, 0, sizeof(to));
    to.td = out->td ? out->td : pl->d_td.p; to.x_td = out->x_td; to.nF = pl->nF;
    to.want_block = out->td != nullptr; to.want_kurt = (out->td != nullptr) || d.has_ku;
    // Default flags consume one bit of the TD features per frame (crest > td_gate_threshold): the float32 filter
    // decides it wherever the crest factor is outside a guard band around the threshold, and the float64 filter
    // re-decides the tiles holding a frame inside the band -- the gate plane equals the float64 one bit for bit.
    const bool fast_td = pl->td_fast && !out->td && !out->x_td && !d.has_ku && d.gate_thr != 0.0f && !pl->td_blocks;

    // start of the run on the caller's stream: selection state and histograms, then the fork
    if (!d.suppressor_bypass) {
        CUDA_OK(ctx, cudaMemsetAsync(hist + (size_t)clip0 * 2 * SEL_BINS, 0, hist_bytes, st));
        select_init_kernel<<<(n_clips + 127) / 128, 128, 0, st>>>(b, d.K, pl->d_sel.p);
        pl->last_launches++;
    }
    if (out->td_fast_crest)    // diagnostic plane: frames of tiles the float32 kernel leaves to the exact one stay 0
        CUDA_OK(ctx, cudaMemsetAsync(out->td_fast_crest + pl->frame_off[clip0], 0,
                                     sizeof(float) * (size_t)(pl->frame_off[clip0 + n_clips] - pl->frame_off[clip0]), st));
    const bool tracing = piped && pl->trace;
    if (tracing) { CUDA_OK(ctx, cudaEventRecord(pl->tr_origin, st)); pl->tr_nseg = n_seg; }
    if (piped) {
        CUDA_OK(ctx, cudaEventRecord(pl->ev_fork, st));
        for (int k = 0; k < apt_plan::SK_N; k++) CUDA_OK(ctx, cudaStreamWaitEvent(S[k], pl->ev_fork, 0));
    }
    // trace marks: j = 0 before the launch on its stream (after its waits), j = 1 after it
    auto tmark = [&](int kind, int sg, int j) { if (tracing) cudaEventRecord(pl->tr_ev[kind][sg][j], S[kind]); };
    int rc = 0;
    cudaError_t e = cudaSuccess;
    const char* what = "";
#define RR(...) do { if (e == cudaSuccess) { e = (__VA_ARGS__); if (e != cudaSuccess) what = #__VA_ARGS__; } } while (0)
    // two-phase schedule: every bulk launch (STFT, TD of all segments) is enqueued before the first chain launch, and the
    // chain waits for the last of them, so that the bulk kernels have the GPU to themselves
    // two_phase == 2 ("three-phase"): the STFT of every segment, then the TD kernel of every segment on the same stream,
    // and the chain starts behind the last STFT -- its first half (tracker pass 1, flux, baselines) needs nothing from the
    // TD kernel, so those latency-bound kernels run beside it; decide waits for its segment's TD launch as always
    const int n_phase = !(two_phase && piped) ? 1 : (two_phase == 2 ? 3 : 2);
    for (int phase = 0; phase < n_phase; phase++)
    for (int sg = 0; sg < n_seg && e == cudaSuccess; sg++) {
        const bool do_stft = n_phase == 1 || phase == 0;
        const bool do_td = n_phase == 1 || (n_phase == 2 ? phase == 0 : phase == 1);
        const bool do_chain = n_phase == 1 || phase == n_phase - 1;
        Batch bs = b;
        bs.ta = sg * seg_frames;
        bs.tb = n_seg == 1 ? INT32_MAX : bs.ta + seg_frames;
        if (do_stft) {
        // STFT
        pl->mark(APT_KERNEL_STFT, st);
        bs.tile0 = bs.ta / STFT_TF;
        bs.ntile = n_seg == 1 ? INT32_MAX : seg_frames / STFT_TF;
        {
            const dim3 g = persistent_grid(pl, seg_grid(pl->stft_tile_off, clip0, n_clips, bs.tile0, n_seg == 1 ? INT64_MAX : seg_frames / STFT_TF), 3);
            tmark(apt_plan::SK_STFT, sg, 0);
            if (pl->generic) {
                bs.tile0 = 0;
                RR(pl->prm.fft_f64 ? launch_stft_generic<double, PCM>(pl, bs, pcm, so, S[apt_plan::SK_STFT])
                                   : launch_stft_generic<float, PCM>(pl, bs, pcm, so, S[apt_plan::SK_STFT]));
            } else
            RR(launch_stft_mode<PCM>(pl, bs, g, pcm, so, S[apt_plan::SK_STFT]));
            RR(rec(apt_plan::SK_STFT, sg));
        }
        }
        if (do_td) {
        // TD features
        pl->mark(APT_KERNEL_TD, st);
        bs.tile0 = bs.ta / TD_FT;
        bs.ntile = n_seg == 1 ? INT32_MAX : seg_frames / TD_FT;
        {
            const dim3 g = persistent_grid(pl, seg_grid(pl->td_tile_off, clip0, n_clips, bs.tile0, n_seg == 1 ? INT64_MAX : seg_frames / TD_FT), fast_td ? 3 : 2);
            tmark(apt_plan::SK_TD, sg, 0);
            if (fast_td) {
                TdOut tf = to;
                tf.td = nullptr; tf.want_block = 0; tf.want_kurt = 0;
                tf.gate = pl->d_gate.p; tf.crest_dbg = out->td_fast_crest; tf.guard = std::max(pl->td_guard, 1e-3f);
                // flagged tiles of this segment: its own slice of the list (at most every tile of the segment) and counter
                int64_t max_tiles = 1;
                for (int c = clip0; c < clip0 + n_clips; c++) max_tiles = std::max(max_tiles, pl->td_tile_off[c + 1] - pl->td_tile_off[c]);
                const int64_t per_seg = (n_seg == 1 ? max_tiles : (int64_t)(seg_frames / TD_FT)) * n_clips;
                const int64_t off = std::min<int64_t>((int64_t)sg * per_seg, pl->td_list_cap);
                tf.list = pl->d_td_list.p + off;
                tf.list_cap = (int)std::min<int64_t>(per_seg, pl->td_list_cap - off);
                tf.list_count = pl->d_td_cnt.p + sg;
                RR(cudaMemsetAsync(tf.list_count, 0, sizeof(int), S[apt_plan::SK_TD]));
                if (pl->tdf_ok && pl->td_ns == 2) {
                    TdFastParams q = pl->tdf;
                    q.guard = pl->td_guard; q.gate = tf.gate; q.crest_dbg = tf.crest_dbg;
                    q.list = tf.list; q.list_count = tf.list_count; q.list_cap = tf.list_cap;
                    const dim3 gt = seg_grid(pl->td_tile_off, clip0, n_clips, bs.tile0, n_seg == 1 ? INT64_MAX : seg_frames / TD_FT);
                    if (gt.x > 0 && e == cudaSuccess) {
                        auto kern = td_gate_fast_kernel<PCM>;
                        RR(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tdf_smem_bytes()));
                        if (e == cudaSuccess) {
                            kern<<<gt, TDF_NT, tdf_smem_bytes(), S[apt_plan::SK_TD]>>>(bs, pcm, pl->d_td_tile_off.p, q);
                            pl->last_launches++;
                            RR(cudaGetLastError());
                        }
                    }
                } else {
                    RR(launch_td<PCM, float>(pl, bs, g, pcm, tf, S[apt_plan::SK_TD]));
                }
                if (g.x > 0 && tf.list_cap > 0) RR(launch_td_recheck<PCM>(pl, bs, pcm, tf, S[apt_plan::SK_TD]));
            } else if (pl->td_blocks) {
                // block statistics of the prefiltered waveform, then the crest factor of every (n_fft, hop) frame from them
                TdOut tg = to;
                tg.td = nullptr; tg.want_block = 0; tg.want_kurt = 0;
                tg.blk_sum = pl->d_blk_sum.p; tg.blk_max = pl->d_blk_max.p;
                RR(launch_td<PCM, double>(pl, bs, g, pcm, tg, S[apt_plan::SK_TD]));
                if (e == cudaSuccess) {
                    CrestIO cio{pl->d_blk_sum.p, pl->d_blk_max.p, to.td, pl->nF, out->td != nullptr};
                    crest_blocks_kernel<<<dim3((unsigned)((maxT + 255) / 256), (unsigned)n_clips), 256, 0, S[apt_plan::SK_TD]>>>(pl->dp, bs, cio);
                    pl->last_launches++;
                    RR(cudaGetLastError());
                }
            } else {
                RR(launch_td<PCM, double>(pl, bs, g, pcm, to, S[apt_plan::SK_TD]));
            }
            RR(rec(apt_plan::SK_TD, sg));
        }
        }
        if (!do_chain) continue;
        if (two_phase && piped && sg == 0) {
            if (n_phase == 2) {
                RR(cudaStreamWaitEvent(S[apt_plan::SK_TRK1], pl->ev_seg[apt_plan::SK_TD][n_seg - 1], 0));
                RR(cudaStreamWaitEvent(S[apt_plan::SK_FLUX], pl->ev_seg[apt_plan::SK_TD][n_seg - 1], 0));
            }
            RR(cudaStreamWaitEvent(S[apt_plan::SK_FLUX], pl->ev_seg[apt_plan::SK_STFT][n_seg - 1], 0));
            RR(cudaStreamWaitEvent(S[apt_plan::SK_TRK1], pl->ev_seg[apt_plan::SK_STFT][n_seg - 1], 0));
        }
        // tracker pass 1 on the mode bins (every band bin when one of its planes is requested)
        pl->mark(APT_KERNEL_TRK1, st);
        if (PS) {
            const int64_t lanes = (int64_t)n_clips * d.K;
            RR(wait(apt_plan::SK_TRK1, apt_plan::SK_STFT, sg));
            if (e == cudaSuccess) {
                presmooth_kernel<<<(unsigned)((lanes + 127) / 128), 128, 0, S[apt_plan::SK_TRK1]>>>(pl->dp, bs, pl->d_Pband.p, pl->d_Psm.p,
                                                                                                   pl->d_st_ps.p, pl->st_stride);
                pl->last_launches++;
                RR(cudaGetLastError());
            }
            if (!d.use_norm) RR(rec(apt_plan::SK_TRK1, sg));
        }
        if (d.use_norm) {
            Trk1IO io;
            io.P_band = trk_in; io.NL = nl_plane; io.det_noise_psd = n1_raw; io.nF = pl->nF;
            io.state = pl->d_st_trk1.p; io.state_stride = pl->st_stride;
            const int64_t lanes = (int64_t)n_clips * tab.n_lanes;
            RR(wait(apt_plan::SK_TRK1, apt_plan::SK_STFT, sg));
            tmark(apt_plan::SK_TRK1, sg, 0);
            if (e == cudaSuccess) {
                trk1_kernel<<<(unsigned)((lanes + 127) / 128), 128, 0, S[apt_plan::SK_TRK1]>>>(pl->dp, bs, tab, io);
                pl->last_launches++;
                RR(cudaGetLastError());
            }
            if (post1 && e == cudaSuccess) {
                const dim3 g = frame_lane_grid(bs, tab.n_lanes);
                if (g.x > 0) {
                    if (MED) { median_time_kernel<<<g, 256, 0, S[apt_plan::SK_TRK1]>>>(pl->dp, bs, tab, n1_raw, n1_fin); pl->last_launches++; }
                    lag_clamp_kernel<<<g, 256, 0, S[apt_plan::SK_TRK1]>>>(pl->dp, bs, tab, n1_fin, pl->d_Pband.p, nl_plane, tab.nls);
                    pl->last_launches++;
                    RR(cudaGetLastError());
                }
            }
            RR(rec(apt_plan::SK_TRK1, sg));
        }
        // dB normalisation, flux, per-mode sums
        pl->mark(APT_KERNEL_FLUX, st);
        {
            FluxIO io;
            io.P_band = pl->d_Pband.p; io.NL = nl_plane; io.nls = tab.nls; io.mf = pl->d_mf.p; io.stride = pl->mf_stride;
            io.det_noise_lag = out->det_noise_lag; io.D = D_plane; io.mode_flux = out->mode_flux; io.nF = pl->nF;
            io.ft = pl->flux_ft;
            bs.tile0 = bs.ta / pl->flux_ft;
            const dim3 g = seg_grid(pl->flux_tile_off, clip0, n_clips, bs.tile0, n_seg == 1 ? INT64_MAX : seg_frames / pl->flux_ft);
            const size_t fsm = flux_smem_bytes(pl->flux_ft, tab.n_lanes);
            RR(wait(apt_plan::SK_FLUX, d.use_norm ? apt_plan::SK_TRK1 : apt_plan::SK_STFT, sg));
            tmark(apt_plan::SK_FLUX, sg, 0);
            if (g.x > 0 && e == cudaSuccess) {
                RR(cudaFuncSetAttribute(flux_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
                if (e == cudaSuccess) {
                    flux_kernel<<<g, 256, fsm, S[apt_plan::SK_FLUX]>>>(pl->dp, bs, pl->d_flux_tile_off.p, tab, io);
                    pl->last_launches++;
                    RR(cudaGetLastError());
                }
            }
            RR(rec(apt_plan::SK_FLUX, sg));
        }
        // float64 baselines + normalisation
        pl->mark(APT_KERNEL_BASE, st);
        {
            const int64_t lanes = (int64_t)n_clips * (d.M + 1);
            RR(wait(apt_plan::SK_BASE, apt_plan::SK_FLUX, sg));
            tmark(apt_plan::SK_BASE, sg, 0);
            if (e == cudaSuccess) {
                base_kernel<<<(unsigned)((lanes + 127) / 128), 128, 0, S[apt_plan::SK_BASE]>>>(pl->dp, bs, pl->d_mf.p, pl->mf_stride, pl->d_st_base.p);
                pl->last_launches++;
                RR(cudaGetLastError());
            }
            RR(rec(apt_plan::SK_BASE, sg));
        }
        // decision
        pl->mark(APT_KERNEL_DECIDE, st);
        {
            DecIO io;
            io.mf = pl->d_mf.p; io.stride = pl->mf_stride; io.td = to.td; io.gate_in = fast_td ? pl->d_gate.p : nullptr;
            io.frame_class = out->frame_class; io.rain_conf = out->rain_conf; io.noise_conf = out->noise_conf;
            io.norm_flux = out->norm_flux; io.score = out->score; io.gate = out->gate; io.nF = pl->nF;
            const int64_t fr = std::min<int64_t>(maxT - bs.ta, n_seg == 1 ? maxT : seg_frames);
            RR(wait(apt_plan::SK_DEC, apt_plan::SK_BASE, sg));
            RR(wait(apt_plan::SK_DEC, apt_plan::SK_TD, sg));
            tmark(apt_plan::SK_DEC, sg, 0);
            if (fr > 0 && e == cudaSuccess) {
                decide_kernel<<<dim3((unsigned)((fr + 255) / 256), (unsigned)n_clips), 256, 0, S[apt_plan::SK_DEC]>>>(pl->dp, bs, io);
                pl->last_launches++;
                RR(cudaGetLastError());
            }
            RR(rec(apt_plan::SK_DEC, sg));
        }
        if (!d.suppressor_bypass) {
            // tracker pass 2, then the noise-floor dB sums and histogram of the segment
            pl->mark(APT_KERNEL_TRK2, st);
            Trk2IO io;
            io.P_band = trk_in; io.frame_class = out->frame_class; io.N2 = n2_raw; io.nF = pl->nF;
            io.state = pl->d_st_trk2.p; io.state_stride = pl->st_stride; io.aq_state = pl->d_st_aq.p;
            const int64_t lanes = (int64_t)n_clips * d.K;
            // Every lane runs the whole time chain, so the kernel's time is (waves of CTAs) x (chain time at that
            // residency); measured per frame step: ~170 cycles up to 4 warps/SM, 185 at 8, 228 at 12, and a cliff
            // beyond (profiles/r1).  The default is 128-thread CTAs, three per SM (registers).
            // When the launch holds 12..16 warps per SM that leaves a short second wave after a full first one;
            // two even waves of one 256-thread CTA per SM (unused dynamic shared memory as the occupancy limiter)
            // are faster there when the kernel has the GPU to itself (1 000 x 71 lanes: 10.8 -> 9.7 ms).
            const double warps_per_sm = (double)((lanes + 31) / 32) / (double)std::max(1, ctx->sm_count);
            RR(wait(apt_plan::SK_TRK2, apt_plan::SK_DEC, sg));
            if (PS && !d.use_norm) RR(wait(apt_plan::SK_TRK2, apt_plan::SK_TRK1, sg));
            tmark(apt_plan::SK_TRK2, sg, 0);
            if (e == cudaSuccess) {
                if (d.adaptive_q) {
                    trk2_kernel<128, true><<<(unsigned)((lanes + 127) / 128), 128, 0, S[apt_plan::SK_TRK2]>>>(pl->dp, bs, io);
                } else if (!piped && warps_per_sm > 12.0 && warps_per_sm <= 16.0) {
                    RR(cudaFuncSetAttribute(trk2_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
                    if (e == cudaSuccess) trk2_kernel<256><<<(unsigned)((lanes + 255) / 256), 256, 120 * 1024, S[apt_plan::SK_TRK2]>>>(pl->dp, bs, io);
                } else {
                    trk2_kernel<128><<<(unsigned)((lanes + 127) / 128), 128, 0, S[apt_plan::SK_TRK2]>>>(pl->dp, bs, io);
                }
                pl->last_launches++;
                RR(cudaGetLastError());
                if (MED && e == cudaSuccess) {
                    const dim3 g = frame_lane_grid(bs, pl->tab_all.n_lanes);
                    if (g.x > 0) {
                        median_time_kernel<<<g, 256, 0, S[apt_plan::SK_TRK2]>>>(pl->dp, bs, pl->tab_all, n2_raw, n2_plane);
                        pl->last_launches++;
                        RR(cudaGetLastError());
                    }
                }
            }
            RR(rec(apt_plan::SK_TRK2, sg));
            pl->mark(APT_KERNEL_DB, st);
            bs.tile0 = bs.ta / DB_CF;
            const dim3 g = seg_grid(pl->sel_chunk_off, clip0, n_clips, bs.tile0, n_seg == 1 ? INT64_MAX : seg_frames / DB_CF);
            RR(wait(apt_plan::SK_DBS, apt_plan::SK_TRK2, sg));
            tmark(apt_plan::SK_DBS, sg, 0);
            if (g.x > 0 && e == cudaSuccess) {
                dbsum_kernel<<<g, 256, 0, S[apt_plan::SK_DBS]>>>(pl->dp, bs, n2_plane, pl->d_sel_chunk_off.p, hist, pl->d_dbsum.p);
                pl->last_launches++;
                RR(cudaGetLastError());
            }
            RR(rec(apt_plan::SK_DBS, sg));
        }
    }
    // join: the last event of the last kind covers everything before it (each kind's stream is in segment order and
    // waits for its producer kind)
    if (piped) {
        const int last_kind = d.suppressor_bypass ? apt_plan::SK_DEC : apt_plan::SK_DBS;
        cudaError_t ej = cudaStreamWaitEvent(st, pl->ev_seg[last_kind][n_seg - 1], 0);
        if (e != cudaSuccess || ej != cudaSuccess) {
            // something failed after the fork: drain the side streams before reporting, so that nothing of this call is
            // still writing the caller's buffers when the error reaches it
            for (int k = 0; k < apt_plan::SK_N; k++) cudaStreamSynchronize(S[k]);
            if (e == cudaSuccess) { e = ej; what = "cudaStreamWaitEvent(join)"; }
        }
    }
    if (e != cudaSuccess) return fail(ctx, -11, "%s failed: %s", what, cudaGetErrorString(e));
#undef RR
    (void)rc;
    // ---- tail on the caller's stream, whole clips
    Batch bw = b;
    if (want_peaks) {
        PeakIO pio;
        pio.D = D_plane; pio.ratio = out->peak_ratio; pio.gate_score = out->peak_gate_score;
        pio.valid_count = out->peak_valid_count; pio.count_by_mode = out->peak_count_by_mode; pio.nF = pl->nF;
        peak_kernel<<<dim3((unsigned)((maxT + 127) / 128), (unsigned)n_clips), 128, 0, st>>>(pl->dp, bw, pio);
        pl->last_launches++;
        CUDA_OK(ctx, cudaGetLastError());
    }
    pl->mark(APT_KERNEL_DECIDE, st);
    compact_kernel<<<(n_clips * 32 + 127) / 128, 128, 0, st>>>(bw, out->frame_class, out->event_idx, out->event_count);
    pl->last_launches++;
    CUDA_OK(ctx, cudaGetLastError());
    if (!d.suppressor_bypass) {
        if (want_gain) {
            pl->mark(APT_KERNEL_GAIN, st);
            GainIO gio;
            gio.P_band = pl->d_Pband.p; gio.N2 = n2_plane; gio.frame_class = out->frame_class; gio.G = out->G;
            gio.ratio_med = out->ratio_med; gio.nF = pl->nF;
            gio.snr_mode = (out->snr_mode && out->snr_gate) ? out->snr_mode : nullptr; gio.snr_gate = gio.snr_mode ? out->snr_gate : nullptr;
            gain_kernel<<<seg_grid(pl->flux_tile_off, clip0, n_clips, 0, INT64_MAX), 256, 0, st>>>(pl->dp, bw, pl->d_flux_tile_off.p, gio);
            const int64_t lanes = (int64_t)n_clips * d.K;
            gain_time_kernel<<<(unsigned)((lanes + 127) / 128), 128, 0, st>>>(pl->dp, bw, out->frame_class, out->G);
            pl->last_launches += 2;
            if (out->S_hat) {
                const int64_t fbeg = pl->frame_off[clip0], fend = pl->frame_off[clip0 + n_clips];
                const int64_t n = (fend - fbeg) * d.F;
                shat_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pl->dp, fbeg, fend, out->G, out->S, out->S_hat);
                pl->last_launches++;
            }
            if (out->y) {
                int64_t max_n = 1;
                for (int c = clip0; c < clip0 + n_clips; c++) max_n = std::max(max_n, pl->len[c]);
                const dim3 grid((unsigned)((max_n + ISTFT_TH * 128 - 1) / (ISTFT_TH * 128)), (unsigned)n_clips);
                istft256_kernel<<<grid, ISTFT_NT, 0, st>>>(pl->dp, bw, out->S_hat, pl->d_win64.p, pl->d_tw256_64.p, out->y);
                pl->last_launches++;
            }
            CUDA_OK(ctx, cudaGetLastError());
        }
        // exact median: bins of the two middle ranks, candidates of those bins, select inside the candidates;
        // clips whose candidates overflow go through the full 3-level radix select (no-ops for the others)
        pl->mark(APT_KERNEL_SELECT, st);
        const dim3 gall = seg_grid(pl->sel_chunk_off, clip0, n_clips, 0, INT64_MAX);
        sel_scan0_kernel<<<(n_clips * 32 + 127) / 128, 128, 0, st>>>(clip0, n_clips, d.eps32, pl->d_sel.p, hist);
        sel_collect_kernel<<<gall, 256, 0, st>>>(pl->dp, bw, n2_plane, pl->d_sel_chunk_off.p, pl->d_sel.p, pl->d_cand_off.p, pl->d_cand.p);
        sel_final_kernel<<<n_clips, 256, 0, st>>>(clip0, pl->d_sel.p, pl->d_cand_off.p, pl->d_cand.p);
        pl->last_launches += 3;
        CUDA_OK(ctx, cudaGetLastError());
        for (int level = 0; level < 3; level++) {
            CUDA_OK(ctx, cudaMemsetAsync(hist + (size_t)clip0 * 2 * SEL_BINS, 0, hist_bytes, st));
            select_hist_kernel<<<gall, 256, 0, st>>>(pl->dp, bw, n2_plane, pl->d_sel_chunk_off.p, level, pl->d_sel.p, hist);
            select_scan_kernel<<<(n_clips * 32 + 127) / 128, 128, 0, st>>>(clip0, n_clips, level, pl->d_sel.p, hist);
            pl->last_launches += 2;
        }
        CUDA_OK(ctx, cudaGetLastError());
    }
    pl->mark(APT_KERNEL_FINALIZE, st);
    finalize_kernel<<<(n_clips + 127) / 128, 128, 0, st>>>(pl->dp, bw, pl->d_sel.p, pl->d_dbsum.p, pl->d_sel_chunk_off.p, out->event_count, out->clip_stats, 0);
    pl->last_launches++;
    CUDA_OK(ctx, cudaGetLastError());
    pl->mark(-1, st);
    if (tracing) cudaEventRecord(pl->tr_end, st);
    return 0;
}

extern "C" {

int apt_run_i16(apt_plan_t* plan, int stages, const int16_t* dev_pcm, const apt_out_t* out, void* stream) {
    if (!plan) return -1;
    if (!dev_pcm || !out) return fail(plan->ctx, -1, "apt_run_i16: null buffer");
    CUDA_OK(plan->ctx, cudaSetDevice(plan->ctx->device));
    plan->last_launches = 0;
    return run_range<int16_t>(plan, stages, 0, plan->n_clips, dev_pcm, out, (cudaStream_t)stream);
}

int apt_run_f32(apt_plan_t* plan, int stages, const float* dev_pcm, const apt_out_t* out, void* stream) {
    if (!plan) return -1;
    if (!dev_pcm || !out) return fail(plan->ctx, -1, "apt_run_f32: null buffer");
    CUDA_OK(plan->ctx, cudaSetDevice(plan->ctx->device));
    plan->last_launches = 0;
    return run_range<float>(plan, stages, 0, plan->n_clips, dev_pcm, out, (cudaStream_t)stream);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// End-to-end host paths.  Clip groups: group g+1 travels host->device on the copy stream while earlier
// groups compute; each group's results go back on its own compute stream (the device->host copy engine
// is separate from the host->device one).  `feed(g, c0, c1, dst_dev, bytes)` enqueues the host->device
// copy of group g on s_copy (and does whatever host-side staging it needs first).
// ---------------------------------------------------------------------------------------------
struct HostOut {
    int8_t* frame_class; float* rain_conf; float* noise_conf; int32_t* event_idx; int32_t* event_count; float* clip_stats;
};

static int host_groups(const apt_plan* pl) {
    int want_groups = 24;
    if (const char* e = getenv("APT_HOST_GROUPS")) want_groups = std::max(1, atoi(e));
    return std::min(pl->n_clips, want_groups);
}

template <typename PCM, typename Feed>
static int run_host_impl(apt_plan* pl, PCM* d_pcm, Feed feed, const HostOut& ho) {
    apt_ctx* ctx = pl->ctx;
    if (!pl->d_fc.p) {
        CUDA_OK(ctx, pl->d_fc.alloc((size_t)pl->nF)); CUDA_OK(ctx, pl->d_rc.alloc((size_t)pl->nF)); CUDA_OK(ctx, pl->d_nc.alloc((size_t)pl->nF));
        CUDA_OK(ctx, pl->d_ev.alloc((size_t)pl->nF)); CUDA_OK(ctx, pl->d_evc.alloc((size_t)pl->n_clips));
        CUDA_OK(ctx, pl->d_stats.alloc((size_t)pl->n_clips * APT_N_CLIP_STATS));
    }
    apt_out_t o;
    memset(&o, 0, sizeof(o));
    o.frame_class = pl->d_fc.p; o.rain_conf = pl->d_rc.p; o.noise_conf = pl->d_nc.p;
    o.event_idx = pl->d_ev.p; o.event_count = pl->d_evc.p; o.clip_stats = pl->d_stats.p;
    const int n_groups = host_groups(pl);
    std::vector<cudaEvent_t> ev(n_groups, nullptr);
    const bool was_timing = pl->timing;
    pl->timing = false;   // per-kernel event marks assume one stream
    int rc = 0;
    cudaError_t ce = cudaSuccess;
    const char* what = "";
#define HP(...) do { if (ce == cudaSuccess) { ce = (__VA_ARGS__); if (ce != cudaSuccess) what = #__VA_ARGS__; } } while (0)
    for (int g = 0; g < n_groups && rc == 0 && ce == cudaSuccess; g++) {
        const int c0 = (int)((int64_t)pl->n_clips * g / n_groups), c1 = (int)((int64_t)pl->n_clips * (g + 1) / n_groups);
        const int64_t s0 = pl->samp_off[c0], s1 = pl->samp_off[c1];
        const int64_t f0 = pl->frame_off[c0], f1 = pl->frame_off[c1];
        cudaStream_t sc = pl->s_comp[g % apt_plan::N_COMP];
        HP(cudaEventCreateWithFlags(&ev[g], cudaEventDisableTiming));
        HP(feed(g, c0, c1, d_pcm + s0, (size_t)(s1 - s0) * sizeof(PCM)));
        HP(cudaEventRecord(ev[g], pl->s_copy));
        HP(cudaStreamWaitEvent(sc, ev[g], 0));
        if (ce != cudaSuccess) break;
        rc = run_range<PCM>(pl, APT_STAGE_FULL, c0, c1 - c0, d_pcm, &o, sc, 8);
        if (rc != 0) break;
        if (ho.frame_class) HP(cudaMemcpyAsync(ho.frame_class + f0, pl->d_fc.p + f0, (size_t)(f1 - f0), cudaMemcpyDeviceToHost, sc));
        if (ho.rain_conf) HP(cudaMemcpyAsync(ho.rain_conf + f0, pl->d_rc.p + f0, (size_t)(f1 - f0) * 4, cudaMemcpyDeviceToHost, sc));
        if (ho.noise_conf) HP(cudaMemcpyAsync(ho.noise_conf + f0, pl->d_nc.p + f0, (size_t)(f1 - f0) * 4, cudaMemcpyDeviceToHost, sc));
        if (ho.event_idx) HP(cudaMemcpyAsync(ho.event_idx + f0, pl->d_ev.p + f0, (size_t)(f1 - f0) * 4, cudaMemcpyDeviceToHost, sc));
        if (ho.event_count) HP(cudaMemcpyAsync(ho.event_count + c0, pl->d_evc.p + c0, (size_t)(c1 - c0) * 4, cudaMemcpyDeviceToHost, sc));
        if (ho.clip_stats) HP(cudaMemcpyAsync(ho.clip_stats + (size_t)c0 * APT_N_CLIP_STATS, pl->d_stats.p + (size_t)c0 * APT_N_CLIP_STATS,
                                              (size_t)(c1 - c0) * APT_N_CLIP_STATS * 4, cudaMemcpyDeviceToHost, sc));
    }
#undef HP
    // always drain every stream this call touched, whatever failed above: the caller's buffers must not be written
    // to after we return
    cudaError_t e0 = cudaStreamSynchronize(pl->s_copy), e1 = cudaSuccess;
    for (int i = 0; i < apt_plan::N_COMP; i++) { cudaError_t e = cudaStreamSynchronize(pl->s_comp[i]); if (e != cudaSuccess) e1 = e; }
    for (int k = 0; k < apt_plan::SK_N; k++) if (pl->s_kind[k]) { cudaError_t e = cudaStreamSynchronize(pl->s_kind[k]); if (e != cudaSuccess) e1 = e; }
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    pl->timing = was_timing;
    if (rc != 0) return rc;
    if (ce != cudaSuccess) return fail(ctx, -12, "host path: %s failed: %s", what, cudaGetErrorString(ce));
    if (e0 != cudaSuccess) return fail(ctx, -12, "copy stream: %s", cudaGetErrorString(e0));
    if (e1 != cudaSuccess) return fail(ctx, -12, "compute stream: %s", cudaGetErrorString(e1));
    return 0;
}

// Staging of pageable clips into the pinned ring: `n_thr` helper threads copy each group's bytes in equal shares
// (a share may span several clips), one group at a time, released by the main thread (which waits for the slot's
// previous host->device copy first).
struct StagePool {
    const apt_plan* pl;
    const void* const* clips;
    size_t esz;
    int n_groups, n_thr;
    std::mutex mu;
    std::condition_variable cv;
    int go = 0;                       // groups released so far
    std::vector<int> done;            // helper threads finished with group g
    std::vector<std::thread> th;
    bool stop = false;

    void copy_share(int g, int tid) {
        const int n_clips = pl->n_clips;
        const int c0 = (int)((int64_t)n_clips * g / n_groups), c1 = (int)((int64_t)n_clips * (g + 1) / n_groups);
        const int64_t s0 = pl->samp_off[c0], total = pl->samp_off[c1] - s0;
        // shares are cut on 4 KiB boundaries of the group's byte range
        const int64_t bytes = total * (int64_t)esz;
        const int64_t pages = (bytes + 4095) / 4096;
        const int64_t b0 = std::min(bytes, pages * tid / n_thr * 4096), b1 = std::min(bytes, pages * (tid + 1) / n_thr * 4096);
        if (b1 <= b0) return;
        char* dst = (char*)pl->ring[g % apt_plan::N_RING];
        // first clip whose byte range reaches b0
        int c = (int)(std::upper_bound(pl->samp_off.begin() + c0, pl->samp_off.begin() + c1 + 1, s0 + b0 / (int64_t)esz) - pl->samp_off.begin()) - 1;
        int64_t pos = b0;
        while (pos < b1 && c < c1) {
            const int64_t cb0 = (pl->samp_off[c] - s0) * (int64_t)esz, cb1 = (pl->samp_off[c + 1] - s0) * (int64_t)esz;
            const int64_t e = std::min(b1, cb1);
            if (e > pos) memcpy(dst + pos, (const char*)clips[c] + (pos - cb0), (size_t)(e - pos));
            pos = e;
            c++;
        }
    }
    std::vector<char> skipped;        // groups that need no staging
    void skip(int g) {
        {
            std::lock_guard<std::mutex> lk(mu);
            skipped[g] = 1;
            go = g + 1;
        }
        cv.notify_all();
    }
    void worker(int tid) {
        for (int g = 0; g < n_groups; g++) {
            bool sk;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return go > g || stop; });
                if (stop) return;
                sk = skipped[g] != 0;
            }
            if (!sk) copy_share(g, tid);
            {
                std::lock_guard<std::mutex> lk(mu);
                done[g]++;
            }
            cv.notify_all();
        }
    }
    void start() {
        done.assign(n_groups, 0);
        skipped.assign(n_groups, 0);
        for (int i = 0; i < n_thr; i++) th.emplace_back([this, i] { worker(i); });
    }
    void release_and_wait(int g) {
        {
            std::lock_guard<std::mutex> lk(mu);
            go = g + 1;
        }
        cv.notify_all();
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return done[g] == n_thr; });
    }
    void finish() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        for (auto& t : th) t.join();
        th.clear();
    }
};

template <typename PCM>
static int run_host_clips_impl(apt_plan* pl, const void* const* clips, const HostOut& ho) {
    apt_ctx* ctx = pl->ctx;
    const int n_groups = host_groups(pl);
    size_t max_bytes = 0;
    for (int g = 0; g < n_groups; g++) {
        const int c0 = (int)((int64_t)pl->n_clips * g / n_groups), c1 = (int)((int64_t)pl->n_clips * (g + 1) / n_groups);
        max_bytes = std::max(max_bytes, (size_t)(pl->samp_off[c1] - pl->samp_off[c0]) * sizeof(PCM));
    }
    PCM* d_pcm;
    if constexpr (sizeof(PCM) == 2) { if (!pl->d_pcm.p) CUDA_OK(ctx, pl->d_pcm.alloc((size_t)pl->nS)); d_pcm = pl->d_pcm.p; }
    else { if (!pl->d_pcm_f32.p) CUDA_OK(ctx, pl->d_pcm_f32.alloc((size_t)pl->nS)); d_pcm = pl->d_pcm_f32.p; }
    // staging is bound by the host's memory bandwidth, not by cores: every hardware thread helps up to ~24
    // (16-vCPU box: 8 threads 437 ms, 16 threads 351 ms, 24 threads 340 ms for 13.4 GB; profiles/r2/staging.txt)
    int n_thr = (int)std::thread::hardware_concurrency();
    n_thr = std::max(1, std::min(24, n_thr));
    if (const char* e = getenv("APT_STAGE_THREADS")) n_thr = std::max(1, std::min(64, atoi(e)));
    StagePool pool;
    pool.pl = pl; pool.clips = clips; pool.esz = sizeof(PCM); pool.n_groups = n_groups; pool.n_thr = n_thr;
    std::vector<cudaEvent_t> slot_free(n_groups, nullptr);   // recorded after the host->device copy of group g
    // a group whose clips already sit back to back in page-locked memory (a loader that filled one pinned buffer, e.g.
    // parse.Mark3BatchLoader) goes to the device straight from there
    auto direct = [&](int c0, int c1) -> bool {
        for (int c = c0; c + 1 < c1; c++)
            if ((const char*)clips[c + 1] != (const char*)clips[c] + (size_t)pl->len[c] * sizeof(PCM)) return false;
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, clips[c0]) != cudaSuccess) { cudaGetLastError(); return false; }
        return at.type == cudaMemoryTypeHost;
    };
    std::vector<char> is_direct(n_groups, 0);
    bool any_staged = false;
    for (int g = 0; g < n_groups; g++) {
        const int c0 = (int)((int64_t)pl->n_clips * g / n_groups), c1 = (int)((int64_t)pl->n_clips * (g + 1) / n_groups);
        is_direct[g] = direct(c0, c1) ? 1 : 0;
        any_staged = any_staged || !is_direct[g];
    }
    if (any_staged) {
        if (pl->ring_bytes < max_bytes) {
            for (int i = 0; i < apt_plan::N_RING; i++) { if (pl->ring[i]) cudaFreeHost(pl->ring[i]); pl->ring[i] = nullptr; }
            pl->ring_bytes = 0;
            for (int i = 0; i < apt_plan::N_RING; i++) CUDA_OK(ctx, cudaHostAlloc(&pl->ring[i], max_bytes, cudaHostAllocDefault));
            pl->ring_bytes = max_bytes;
        }
    }
    auto feed = [&](int g, int c0, int, PCM* dst, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaSuccess;
        if (is_direct[g]) {
            pool.skip(g);
            return cudaMemcpyAsync(dst, clips[c0], bytes, cudaMemcpyHostToDevice, pl->s_copy);
        }
        // the slot's previous copy (the latest staged group with the same slot) must have left the pinned buffer
        for (int j = g - apt_plan::N_RING; j >= 0; j -= apt_plan::N_RING)
            if (slot_free[j]) {
                e = cudaEventSynchronize(slot_free[j]);
                if (e != cudaSuccess) return e;
                break;
            }
        pool.release_and_wait(g);
        e = cudaMemcpyAsync(dst, pl->ring[g % apt_plan::N_RING], bytes, cudaMemcpyHostToDevice, pl->s_copy);
        if (e != cudaSuccess) return e;
        e = cudaEventCreateWithFlags(&slot_free[g], cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
        return cudaEventRecord(slot_free[g], pl->s_copy);
    };
    pool.start();            // no early return between here and finish(): the helper threads must be joined
    const int rc = run_host_impl<PCM>(pl, d_pcm, feed, ho);
    pool.finish();
    for (auto& e : slot_free) if (e) cudaEventDestroy(e);
    return rc;
}

extern "C" {

int apt_run_host_i16(apt_plan_t* pl, const int16_t* host_pcm, int8_t* frame_class, float* rain_conf, float* noise_conf,
                     int32_t* event_idx, int32_t* event_count, float* clip_stats) {
    if (!pl) return -1;
    apt_ctx* ctx = pl->ctx;
    if (!host_pcm) return fail(ctx, -1, "apt_run_host_i16: null PCM");
    if (!pl->full_ok) return fail(ctx, -34, "apt_run_host_i16: the plan's STFT geometry supports the features stage only");
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    pl->last_launches = 0;
    if (!pl->d_pcm.p) CUDA_OK(ctx, pl->d_pcm.alloc((size_t)pl->nS));
    const HostOut ho{frame_class, rain_conf, noise_conf, event_idx, event_count, clip_stats};
    auto feed = [&](int, int c0, int, int16_t* dst, size_t bytes) -> cudaError_t {
        return cudaMemcpyAsync(dst, host_pcm + pl->samp_off[c0], bytes, cudaMemcpyHostToDevice, pl->s_copy);
    };
    return run_host_impl<int16_t>(pl, pl->d_pcm.p, feed, ho);
}

int apt_run_host_clips(apt_plan_t* pl, const void* const* clip_ptrs, int is_f32, int8_t* frame_class, float* rain_conf,
                       float* noise_conf, int32_t* event_idx, int32_t* event_count, float* clip_stats) {
    if (!pl) return -1;
    apt_ctx* ctx = pl->ctx;
    if (!clip_ptrs) return fail(ctx, -1, "apt_run_host_clips: null clip table");
    for (int c = 0; c < pl->n_clips; c++) if (!clip_ptrs[c]) return fail(ctx, -1, "apt_run_host_clips: clip %d is null", c);
    if (!pl->full_ok) return fail(ctx, -34, "apt_run_host_clips: the plan's STFT geometry supports the features stage only");
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    pl->last_launches = 0;
    const HostOut ho{frame_class, rain_conf, noise_conf, event_idx, event_count, clip_stats};
    return is_f32 ? run_host_clips_impl<float>(pl, clip_ptrs, ho) : run_host_clips_impl<int16_t>(pl, clip_ptrs, ho);
}

}  // extern "C"

extern "C" int apt_dsd_run_i16(apt_ctx* ctx, const apt_dsd_params_t* p, int n_clips, const int64_t* clip_len, const double* ts,
                               const int16_t* dev_pcm, double* dev_out, int32_t* dev_n_minutes, int max_minutes, void* stream) {
    if (!ctx) return -1;
    if (!p || !clip_len || !ts || !dev_pcm || !dev_out || !dev_n_minutes || n_clips <= 0 || n_clips > 65535)
        return fail(ctx, -1, "apt_dsd_run_i16: bad arguments");
    const int L = p->frame_length;
    if (L < 64 || L > 4096 || (L & (L - 1)) != 0 || p->hop_length < 1 || p->fs < 1)
        return fail(ctx, -21, "apt_dsd_run_i16: frame_length=%d must be a power of two in 64..4096, hop >= 1", L);
    if (p->apply_window && !p->window) return fail(ctx, -25, "apt_dsd_run_i16: window table missing");
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    DsdDev d;
    memset(&d, 0, sizeof(d));
    d.fs = p->fs; d.L = L; d.hop = p->hop_length; d.window = p->apply_window;
    // index constants exactly as the reference constructor computes them (:33-71)
    const double dF = (double)p->fs / (double)L;
    auto fdiv = [&](double f) { return (int)floor(f / dF); };
    d.n_bins = L / 2;
    d.rain_lo = fdiv(400) + 1; d.rain_hi = fdiv(700);
    d.pft_lo = fdiv(100) + 1; d.pft_hi = fdiv(1500) - 1;
    d.lwin0 = fdiv(300); d.lwin1 = d.lwin0 + 19 - 1;
    d.hwin0 = fdiv(1000); d.hwin1 = d.hwin0 + 19 - 1;
    d.rain_thr = 0.6; d.rain_log_factor = 0.6; d.rain_log_base = 1.13;
    d.max_minutes = max_minutes;
    if (d.pft_hi > d.n_bins || d.rain_hi >= d.n_bins || d.hwin1 >= d.n_bins || d.pft_lo >= d.pft_hi || d.n_bins > 2048)
        return fail(ctx, -22, "apt_dsd_run_i16: fs=%d / frame_length=%d put the analysis bands outside the spectrum", p->fs, L);
    std::vector<int64_t> so(n_clips + 1, 0), fo(n_clips + 1, 0);
    int64_t max_fr = 1;
    for (int c = 0; c < n_clips; c++) {
        const int64_t n = clip_len[c];
        const int64_t nf = n < L ? 0 : (n - L) / p->hop_length + 1;
        if ((int64_t)ceil((double)n / ((double)p->fs * 60.0)) > max_minutes) return fail(ctx, -28, "clip %d needs more than max_minutes=%d rows", c, max_minutes);
        so[c + 1] = so[c] + n; fo[c + 1] = fo[c] + nf;
        max_fr = std::max(max_fr, nf);
    }
    PoolBuf<int64_t> d_so(ctx, 0), d_fo(ctx, 1); PoolBuf<double> d_ts(ctx, 2), d_drop(ctx, 3), d_pkv(ctx, 4), d_win(ctx, 5);
    PoolBuf<int> d_pki(ctx, 6); PoolBuf<cx<double>> d_tw(ctx, 7);
    CUDA_OK(ctx, upload(d_so, so)); CUDA_OK(ctx, upload(d_fo, fo));
    CUDA_OK(ctx, upload(d_ts, std::vector<double>(ts, ts + n_clips)));
    std::vector<cx<double>> tw(L / 2 + 1);
    for (int k = 0; k <= L / 2; k++) tw[k] = {cos(2.0 * M_PI * k / L), -sin(2.0 * M_PI * k / L)};
    CUDA_OK(ctx, upload(d_tw, tw));
    if (p->apply_window) CUDA_OK(ctx, upload(d_win, std::vector<double>(p->window, p->window + L)));
    const size_t nf_tot = (size_t)std::max<int64_t>(1, fo[n_clips]);
    CUDA_OK(ctx, d_drop.alloc(nf_tot)); CUDA_OK(ctx, d_pkv.alloc(nf_tot)); CUDA_OK(ctx, d_pki.alloc(nf_tot));
    const char* e_gen = getenv("APT_DSD_FFT_GENERIC");
    if (L == 512 && !(e_gen && atoi(e_gen) != 0)) {
        // the emulator's frame length: 16 lanes per frame, 16 frames per CTA (APT_DSD_FFT_GENERIC=1 keeps one CTA per frame)
        const size_t smem = dsd_fft512_smem();
        CUDA_OK(ctx, cudaFuncSetAttribute(dsd_fft512_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dsd_fft512_kernel<<<dim3((unsigned)((max_fr + DSD_F512_TF - 1) / DSD_F512_TF), (unsigned)n_clips), DSD_F512_NT, smem, st>>>(
            d, d_so.p, d_fo.p, dev_pcm, d_win.p, d_tw.p, d_drop.p, d_pki.p, d_pkv.p);
    } else {
        const size_t smem = sizeof(cx<double>) * (size_t)L + sizeof(double) * (size_t)(L / 2);
        CUDA_OK(ctx, cudaFuncSetAttribute(dsd_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dsd_frame_kernel<<<dim3((unsigned)max_fr, (unsigned)n_clips), DSD_NT, smem, st>>>(d, 0, d_so.p, d_fo.p, dev_pcm, d_win.p, d_tw.p,
                                                                                           d_drop.p, d_pki.p, d_pkv.p);
    }
    // position-only quantities (timestamps, 2-second slots, drop-size bins) for every hop position, then the serial machine
    PoolBuf<double> d_tsc(ctx, 8); PoolBuf<int> d_aux(ctx, 9);
    CUDA_OK(ctx, d_tsc.alloc(nf_tot + (size_t)n_clips)); CUDA_OK(ctx, d_aux.alloc(nf_tot));
    dsd_times_kernel<<<dim3((unsigned)((max_fr + 1 + 255) / 256), (unsigned)n_clips), 256, 0, st>>>(d, d_fo.p, d_ts.p, d_drop.p, d_tsc.p, d_aux.p);
    const char* e_ser = getenv("APT_DSD_STATE_SERIAL");
    if (e_ser && atoi(e_ser) != 0) {
        dsd_minutes_kernel<<<(n_clips + 63) / 64, 64, 0, st>>>(d, n_clips, d_so.p, d_fo.p, d_ts.p, d_tsc.p, d_aux.p, d_pki.p, d_pkv.p, dev_out, dev_n_minutes);
    } else {
        // one warp per clip, histograms in shared memory (APT_DSD_STATE_SERIAL=1 keeps one thread per clip)
        const size_t smem_m = sizeof(double) * (size_t)(d.n_bins + DSD_OUT) + sizeof(int) * (size_t)d.n_bins;
        dsd_minutes_warp_kernel<<<n_clips, 32, smem_m, st>>>(d, n_clips, d_so.p, d_fo.p, d_ts.p, d_tsc.p, d_aux.p, d_pki.p, d_pkv.p, dev_out, dev_n_minutes);
    }
    CUDA_OK(ctx, cudaGetLastError());
    CUDA_OK(ctx, cudaStreamSynchronize(st));
    return 0;
}

extern "C" int apt_sizeof_bne_params(void) { return (int)sizeof(apt_bne_params_t); }

extern "C" int apt_bne_run(apt_ctx* ctx, const apt_bne_params_t* p, int n_clips, const int64_t* clip_len, const void* dev_pcm, int is_f32,
                           double* dev_frame_out, uint8_t* dev_mask, double* dev_subE, double* dev_stats, void* stream) {
    if (!ctx) return -1;
    if (!p || !clip_len || !dev_pcm || !dev_frame_out || !dev_mask || !dev_subE || !dev_stats || n_clips <= 0 || n_clips > 65535)
        return fail(ctx, -1, "apt_bne_run: bad arguments");
    const int N = p->N;
    if (N < 64 || N > 4096 || (N & (N - 1)) != 0) return fail(ctx, -21, "apt_bne_run: frame_len=%d must be a power of two in 64..4096", N);
    if (p->sub_len < 8 || p->sub_len > 128 || p->sub_len % 8 != 0 || N % p->sub_len != 0 || p->S != N / p->sub_len || p->S > BNE_MAX_S)
        return fail(ctx, -26, "apt_bne_run: subframes of %d samples (need a multiple of 8 up to 128 tiling the frame, at most %d per frame)", p->sub_len, BNE_MAX_S);
    if (p->ns_h < 0 || p->ns_h > BNE_MAX_SOS || p->ns_b < 1 || p->ns_b > BNE_MAX_SOS) return fail(ctx, -24, "apt_bne_run: filter sections out of range");
    if (p->W < 1 || p->W > BNE_MAX_W || p->W_min < 0 || p->W_min > p->W) return fail(ctx, -26, "apt_bne_run: W=%d outside [1,%d]", p->W, BNE_MAX_W);
    if (p->n_bands < 0 || p->n_bands > BNE_MAX_BANDS || p->warm < 0) return fail(ctx, -26, "apt_bne_run: bad band table / warm-up");
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<int64_t> so(n_clips + 1, 0), fo(n_clips + 1, 0), sg(n_clips + 1, 0);
    int64_t max_fr = 1;
    int seg_frames = BNE_SEG;
    if (const char* e = getenv("APT_BNE_SEG")) seg_frames = std::max(1, std::min(4096, atoi(e)));
    for (int c = 0; c < n_clips; c++) {
        const int64_t nf = clip_len[c] / N;
        so[c + 1] = so[c] + clip_len[c]; fo[c + 1] = fo[c] + nf; sg[c + 1] = sg[c] + (nf + seg_frames - 1) / seg_frames;
        max_fr = std::max(max_fr, nf);
    }
    PoolBuf<int64_t> d_so(ctx, 10), d_fo(ctx, 11), d_sg(ctx, 12); PoolBuf<double> d_xhp(ctx, 13), d_subEh(ctx, 14), d_fftq(ctx, 15);
    PoolBuf<cx<double>> d_tw(ctx, 16);
    CUDA_OK(ctx, upload(d_so, so)); CUDA_OK(ctx, upload(d_fo, fo)); CUDA_OK(ctx, upload(d_sg, sg));
    std::vector<cx<double>> tw(N / 2 + 1);
    for (int k = 0; k <= N / 2; k++) tw[k] = {cos(2.0 * M_PI * k / N), -sin(2.0 * M_PI * k / N)};
    CUDA_OK(ctx, upload(d_tw, tw));
    const size_t nf_tot = (size_t)std::max<int64_t>(1, fo[n_clips]);
    CUDA_OK(ctx, d_xhp.alloc((size_t)std::max<int64_t>(1, so[n_clips])));
    CUDA_OK(ctx, d_subEh.alloc(nf_tot * BNE_MAX_S)); CUDA_OK(ctx, d_fftq.alloc(nf_tot * 4));
    const int64_t nseg = sg[n_clips];
    if (nseg > 0) {
        const char* e_ser = getenv("APT_BNE_FILTER_SERIAL");
        const int G = p->ns_h + p->ns_b;
        if (p->ns_h >= 1 && G <= 32 && !(e_ser && atoi(e_ser) != 0)) {
            // wavefront over the cascade: one lane per second-order section (APT_BNE_FILTER_SERIAL=1 keeps one thread per segment)
            const int64_t warps = (nseg + (32 / G) - 1) / (32 / G);
            const unsigned blocks = (unsigned)((warps + 3) / 4);
            const int steps = (int)std::min<int64_t>((int64_t)seg_frames * N + p->warm, max_fr * N) + G - 1;   // longest run + lanes
            if (is_f32) bne_filter_wave_kernel<float><<<blocks, 128, 0, st>>>(*p, n_clips, steps, d_so.p, d_fo.p, d_sg.p, seg_frames, (const float*)dev_pcm, d_xhp.p, d_subEh.p, dev_subE);
            else bne_filter_wave_kernel<int16_t><<<blocks, 128, 0, st>>>(*p, n_clips, steps, d_so.p, d_fo.p, d_sg.p, seg_frames, (const int16_t*)dev_pcm, d_xhp.p, d_subEh.p, dev_subE);
        } else if (is_f32) bne_filter_kernel<float><<<(unsigned)((nseg + 127) / 128), 128, 0, st>>>(*p, n_clips, d_so.p, d_fo.p, d_sg.p, seg_frames, (const float*)dev_pcm, d_xhp.p, d_subEh.p, dev_subE);
        else bne_filter_kernel<int16_t><<<(unsigned)((nseg + 127) / 128), 128, 0, st>>>(*p, n_clips, d_so.p, d_fo.p, d_sg.p, seg_frames, (const int16_t*)dev_pcm, d_xhp.p, d_subEh.p, dev_subE);
        const char* e_gen = getenv("APT_BNE_FFT_GENERIC");
        if (N == 256 && !(e_gen && atoi(e_gen) != 0)) {
            // the sensor's frame length: 8 lanes per frame, 32 frames per CTA (APT_BNE_FFT_GENERIC=1 keeps the generic kernel)
            const size_t smem = bne_fft256_smem();
            CUDA_OK(ctx, cudaFuncSetAttribute(bne_fft256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            bne_fft256_kernel<<<dim3((unsigned)((max_fr + BNE_F256_TF - 1) / BNE_F256_TF), (unsigned)n_clips), BNE_F256_NT, smem, st>>>(
                *p, d_so.p, d_fo.p, d_xhp.p, d_tw.p, d_fftq.p);
        } else if (N == 512 && !(e_gen && atoi(e_gen) != 0)) {
            // the reference's default frame length: 16 lanes per frame, 16 frames per CTA
            const size_t smem = bne_fft512_smem();
            CUDA_OK(ctx, cudaFuncSetAttribute(bne_fft512_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            bne_fft512_kernel<<<dim3((unsigned)((max_fr + BNE_F512_TF - 1) / BNE_F512_TF), (unsigned)n_clips), BNE_F512_NT, smem, st>>>(
                *p, d_so.p, d_fo.p, d_xhp.p, d_tw.p, d_fftq.p);
        } else {
            const size_t smem = sizeof(cx<double>) * (size_t)N + sizeof(double) * (size_t)(N + 2);
            CUDA_OK(ctx, cudaFuncSetAttribute(bne_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            bne_fft_kernel<<<dim3((unsigned)max_fr, (unsigned)n_clips), BNE_NT, smem, st>>>(*p, d_so.p, d_fo.p, d_xhp.p, d_tw.p, d_fftq.p);
        }
    }
    // one warp per clip; APT_BNE_STATE_SERIAL=1 keeps the one-thread-per-clip kernel (bit-equal, for comparison)
    const char* e_serial = getenv("APT_BNE_STATE_SERIAL");
    if (e_serial && atoi(e_serial) != 0)
        bne_state_kernel<<<(n_clips + 63) / 64, 64, 0, st>>>(*p, n_clips, d_fo.p, d_subEh.p, dev_subE, d_fftq.p, dev_frame_out, dev_mask, dev_stats);
    else if (p->S == 4)
        bne_state_warp_kernel<4><<<n_clips, 32, 0, st>>>(*p, n_clips, d_fo.p, d_subEh.p, dev_subE, d_fftq.p, dev_frame_out, dev_mask, dev_stats);
    else
        bne_state_warp_kernel<0><<<n_clips, 32, 0, st>>>(*p, n_clips, d_fo.p, d_subEh.p, dev_subE, d_fftq.p, dev_frame_out, dev_mask, dev_stats);
    CUDA_OK(ctx, cudaGetLastError());
    CUDA_OK(ctx, cudaStreamSynchronize(st));
    return 0;
}

extern "C" int apt_sizeof_roe_params(void) { return (int)sizeof(apt_roe_params_t); }

extern "C" int apt_roe_run(apt_ctx* ctx, const apt_roe_params_t* p, int n_clips, const void* dev_pcm, int is_f32, int n_parts,
                           const int32_t* part_clip, const int64_t* part_start, const int32_t* part_len, int max_harmonics_in,
                           double* dev_frame_out, double* dev_part_out, double* dev_clip_out, int* max_harmonics_out, void* stream) {
    if (!ctx) return -1;
    if (!p || !dev_pcm || !dev_frame_out || !dev_part_out || !dev_clip_out || !max_harmonics_out || n_clips <= 0 || n_parts < 0 ||
        (n_parts > 0 && (!part_clip || !part_start || !part_len)))
        return fail(ctx, -1, "apt_roe_run: bad arguments");
    if (p->n_fft != 256 || p->hop != 128) return fail(ctx, -21, "apt_roe_run: frame %d / hop %d (the CUDA path is built for 256 / 128)", p->n_fft, p->hop);
    if (p->ns_in < 1 || p->ns_in > 8 || p->ns_td < 0 || p->ns_td > 4 || p->ns_in + p->ns_td > 32) return fail(ctx, -24, "apt_roe_run: filter sections out of range");
    if (p->M < 2 || p->wl < 1 || p->wl > 8 || p->max_peaks < 1) return fail(ctx, -26, "apt_roe_run: local-average window out of range (M=%d, wl=%d)", p->M, p->wl);
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<int64_t> fo(n_parts + 1, 0), yo(n_parts + 1, 0);
    std::vector<int> cp0(n_clips + 1, 0);
    int max_T = 1;
    for (int q = 0; q < n_parts; q++) {
        if (part_len[q] < 256 || part_len[q] > 254 * 128) return fail(ctx, -22, "apt_roe_run: part %d has %d samples (256..%d)", q, part_len[q], 254 * 128);
        if (part_clip[q] < 0 || part_clip[q] >= n_clips || (q && part_clip[q] < part_clip[q - 1])) return fail(ctx, -22, "apt_roe_run: parts must be listed clip by clip");
        const int T = 1 + part_len[q] / 128;
        fo[q + 1] = fo[q] + T + 1; yo[q + 1] = yo[q] + part_len[q];
        cp0[part_clip[q] + 1]++;
        max_T = std::max(max_T, T);
    }
    for (int c = 0; c < n_clips; c++) cp0[c + 1] += cp0[c];
    CUDA_OK(ctx, cudaMemsetAsync(dev_clip_out, 0, sizeof(double) * (size_t)n_clips * APT_ROE_CLIP_F, st));
    *max_harmonics_out = max_harmonics_in;
    PoolBuf<int> d_cp0(ctx, 20); CUDA_OK(ctx, upload(d_cp0, cp0));
    if (n_parts > 0) {
        PoolBuf<int64_t> d_fo(ctx, 21), d_yo(ctx, 22), d_start(ctx, 23); PoolBuf<int32_t> d_clip(ctx, 24), d_len(ctx, 25); PoolBuf<int> d_mh(ctx, 26);
        PoolBuf<double> d_y(ctx, 27), d_t(ctx, 28), d_mag(ctx, 29), d_harm(ctx, 30); PoolBuf<cx<double>> d_twA(ctx, 31), d_tw256(ctx, 32);
        CUDA_OK(ctx, upload(d_fo, fo)); CUDA_OK(ctx, upload(d_yo, yo));
        CUDA_OK(ctx, upload(d_start, std::vector<int64_t>(part_start, part_start + n_parts)));
        CUDA_OK(ctx, upload(d_clip, std::vector<int32_t>(part_clip, part_clip + n_parts)));
        CUDA_OK(ctx, upload(d_len, std::vector<int32_t>(part_len, part_len + n_parts)));
        std::vector<cx<double>> twA(128), tw256(129);
        for (int i = 0; i < 128; i++) { const int m = ((i & 7) * (i >> 3)) & 127; twA[i] = {cos(2.0 * M_PI * m / 128.0), -sin(2.0 * M_PI * m / 128.0)}; }
        for (int k = 0; k <= 128; k++) tw256[k] = {cos(2.0 * M_PI * k / 256.0), -sin(2.0 * M_PI * k / 256.0)};
        CUDA_OK(ctx, upload(d_twA, twA)); CUDA_OK(ctx, upload(d_tw256, tw256));
        CUDA_OK(ctx, d_y.alloc((size_t)yo[n_parts])); CUDA_OK(ctx, d_t.alloc((size_t)yo[n_parts] + (size_t)256 * n_parts));
        CUDA_OK(ctx, d_mag.alloc((size_t)(fo[n_parts] - n_parts) * 129)); CUDA_OK(ctx, d_harm.alloc((size_t)fo[n_parts] * 5));
        CUDA_OK(ctx, d_mh.alloc((size_t)n_parts + 1));
        const RoeParts pt{n_parts, d_clip.p, d_start.p, d_len.p, d_fo.p, d_yo.p};
        const char* e_ser = getenv("APT_ROE_FILTER_SERIAL");
        const int nl = p->ns_in + (p->want_td ? p->ns_td : 0);
        if (nl >= 1 && nl <= 32 && !(e_ser && atoi(e_ser) != 0)) {
            // 32 / nl parts per warp, one shuffle per step (APT_ROE_FILTER_SERIAL=1 keeps one warp per part)
            const int warps = (n_parts + (32 / nl) - 1) / (32 / nl);
            const unsigned blocks = (unsigned)((warps + 3) / 4);
            int max_len = 0;
            for (int q = 0; q < n_parts; q++) max_len = std::max(max_len, (int)part_len[q]);
            const int steps = max_len + 128 + nl;
            if (is_f32) roe_filter_wave_kernel<float><<<blocks, 128, 0, st>>>(*p, pt, steps, (const float*)dev_pcm, d_y.p, d_t.p);
            else roe_filter_wave_kernel<int16_t><<<blocks, 128, 0, st>>>(*p, pt, steps, (const int16_t*)dev_pcm, d_y.p, d_t.p);
        } else if (is_f32) roe_filter_kernel<float><<<n_parts, 32, 0, st>>>(*p, pt, (const float*)dev_pcm, d_y.p, d_t.p);
        else roe_filter_kernel<int16_t><<<n_parts, 32, 0, st>>>(*p, pt, (const int16_t*)dev_pcm, d_y.p, d_t.p);
        const size_t smem = sizeof(cx<double>) * (32 * kExSize + 128 + 130) + sizeof(double) * 256;
        CUDA_OK(ctx, cudaFuncSetAttribute(roe_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        roe_frame_kernel<<<dim3((unsigned)((max_T + 31) / 32), (unsigned)n_parts), ROE_NT, smem, st>>>(*p, pt, d_y.p, d_t.p, d_twA.p, d_tw256.p, d_mag.p, dev_frame_out);
        roe_part_kernel<<<n_parts, ROE_NT, 0, st>>>(*p, pt, d_mag.p, dev_frame_out, d_harm.p, dev_part_out);
        roe_state_kernel<<<1, 32, 0, st>>>(pt, dev_part_out, max_harmonics_in, d_mh.p, d_mh.p + n_parts);
        roe_combine_kernel<<<n_parts, ROE_NT, 0, st>>>(*p, pt, d_mh.p, d_harm.p, dev_frame_out, dev_part_out);
        roe_clip_kernel<<<(n_clips + 127) / 128, 128, 0, st>>>(*p, n_clips, d_cp0.p, dev_part_out, dev_clip_out);
        CUDA_OK(ctx, cudaGetLastError());
        CUDA_OK(ctx, cudaMemcpyAsync(max_harmonics_out, d_mh.p + n_parts, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_OK(ctx, cudaStreamSynchronize(st));
    } else {
        CUDA_OK(ctx, cudaStreamSynchronize(st));
    }
    return 0;
}
