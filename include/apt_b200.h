/*
 * apt_b200.h -- C ABI of libapt_b200.so: the B200 (sm_100a) implementation of the
 * audio_processing_tools hot path (framing/windowing -> STFT -> band powers -> noise-PSD tracking
 * -> rain-frame detection -> clip statistics), batched over many clips.
 *
 * The boundary replaces ONE call of the reference:
 *     results, state = proc.run(audio, proc_params)
 *         audio_processing_tools/audio_processing_framework.py:190
 * i.e. everything RainDetectorProcessor.run / SpectralNoiseProcessor.process compute
 *         audio_processing_tools/edge/rain_signal_processor.py:1223-1344, :788-1198
 * for a whole batch of clips at once.  The shape of the ABI follows the reference's own FFI
 * precedent (edge/parameter_tuning/call_c_fun.py:20-58,193-236): POD structs passed by pointer,
 * `int` status return, caller-owned buffers, no library-owned results.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; apt_last_error() describes the last error
 *     of a context.  The library never aborts and never computes on the CPU instead of the GPU.
 *   - one apt_ctx per GPU; a ctx and its plans are not thread-safe.
 *   - apt_run_* enqueue work on the given CUDA stream (cudaStream_t as void*; NULL = default
 *     stream) and return without synchronising.  All *device* buffers are caller-owned
 *     (PyTorch tensors are used only as carriers of such buffers); the plan owns scratch only.
 *   - per-frame arrays of all clips are concatenated; clip c owns frames
 *     [frame_offsets[c], frame_offsets[c+1]) with T_c = 1 + len_c / hop  (librosa center=True,
 *     rain_signal_processor.py:818-825).  Per-sample arrays are concatenated the same way
 *     (sample_offsets).
 */
#ifndef APT_B200_H
#define APT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APT_MAX_MODES 8
#define APT_MAX_SOS 4
#define APT_N_RAW_FEATURES 21 /* feature_extraction.py:9-31 RAW_SPECTRAL_FEATURE_NAMES */
#define APT_N_TD_FEATURES 5   /* crest, kurtosis, block crest, block width50, block post/pre */
#define APT_N_CLIP_STATS 8
#define APT_MAX_PRE_SMOOTH 16
#define APT_MAX_MEDIAN 31
#define APT_MAX_GAIN_TAPS 9
#define APT_ABI_VERSION 10

typedef struct apt_ctx apt_ctx;
typedef struct apt_plan apt_plan_t;

/*
 * Resolved configuration: the POD image of NoiseProcessorConfig + the detector dict
 * (rain_signal_processor.py:19-188, rain_frame_classifier.py:314-459) after the host applied the
 * reference's precedence rules (build_noise_config, :202-255; _dget, rain_frame_classifier.py:135-148)
 * and the casts numpy applies to Python scalars (float32 where the reference multiplies float32
 * arrays by Python floats).  Bin ranges are inclusive rfft bin indices; lo > hi means "empty".
 */
typedef struct apt_params_t {
    int32_t abi_version;
    int32_t fs, n_fft, hop;
    int32_t band_lo, band_hi;                  /* operating band (rain_signal_processor.py:832-833) */
    int32_t n_modes;                           /* >= 4 (rain_frame_classifier.py:379-383) */
    int32_t mode_lo[APT_MAX_MODES], mode_hi[APT_MAX_MODES];            /* over all rfft bins */
    int32_t mode_band_lo[APT_MAX_MODES], mode_band_hi[APT_MAX_MODES];  /* relative to band_lo */
    double  mode_weight[APT_MAX_MODES];        /* 1.0 when mode_weights is None */
    /* noise-PSD tracker, rain_signal_processor.py:555-666 */
    float   trk_eta, trk_scale_alpha, trk_one_minus_alpha, trk_step_floor;
    float   trk_q, trk_neg_one_minus_q, trk_maxr;
    double  ema_up, ema_down;
    int32_t warmup_need;
    float   eps_f32;
    int32_t detector_use_noise_norm;           /* rain_signal_processor.py:862 */
    int32_t norm_ratio_db;                     /* :884 */
    /* flux baseline, rain_frame_classifier.py:31-82 (Python doubles) */
    double  bl_q, bl_eta, bl_scale_alpha, bl_floor;
    int32_t norm_enable;
    float   norm_min_f32;
    /* decision, rain_frame_classifier.py:230-284, :914-998 */
    float   thr_primary, thr_m1, thr_m2, thr_m3;
    int32_t min_support;
    float   td_gate_thr;
    int32_t has_kurt_upper;
    float   kurt_upper;
    float   noise_hi, mode_flux_noise_max;
    /* zero-phase TD prefilter (scipy butter SOS + sosfilt_zi), rain_frame_classifier.py:472-481 */
    int32_t n_sos, padlen;
    double  sos[APT_MAX_SOS][6];
    double  zi[APT_MAX_SOS][2];
    double  eps_f64;
    /* TD block-energy features, feature_extraction.py:253-366 */
    int32_t blk_len, blk_hop, blk_post_pre, blk_smooth;
    /* raw spectral features, feature_extraction.py:542-747 */
    int32_t low_lo, low_hi, rain_lo, rain_hi;
    double  rolloff_fraction;
    int32_t suppressor_bypass;
    int32_t clip_rain_min_frames;              /* rain_signal_processor.py:1256-1257 */
    /* arithmetic of the STFT: 1 = float64 FFT rounded to complex64 (the reference's arithmetic,
       librosa/scipy.fft on float64), 0 = float32 FFT (spectra within 2e-6 of frame max),
       2 = DFT as a GEMM on the tcgen05 tensor cores (n_fft = 256, hop = 128, int16 input, band plane / band energies
       only; exact int8 limbs of the samples x two fp16 limbs of window x twiddle, fp32 accumulation in tensor memory:
       the same tolerance class as 0; requests it cannot serve run as 0) */
    int32_t fft_f64;
    /* suppressor gain, rain_signal_processor.py:400-533 (_compute_gain) and :1028-1091; float32 values are
       the ones numpy forms when the reference mixes Python floats with float32 arrays */
    int32_t gain_mode;                         /* 0 = sqrt_sub (default), 1 = wiener */
    int32_t adaptive_gain, gain_freq_smooth, n_gain_taps, use_lagged_noise_psd;
    float   oversub_noise, oversub_rain;       /* oversubtraction at noise_conf = 1 / 0 (:436-437) */
    float   gain_floor, gain_ceil;
    float   gain_taps[APT_MAX_GAIN_TAPS];      /* normalised frequency-smoothing kernel (:484-487) */
    float   alpha_noise, one_minus_alpha_noise;/* temporal smoothing on noise-like frames (:508-516) */
    float   alpha_base, one_minus_alpha_base;  /* non-adaptive mode (:522-523) */
    float   gain_eps_f32;
    /* optional peak-structure features, rain_frame_classifier.py:670-683, :761-843 (debug outputs only) */
    int32_t peak_top_p, primary_top_m;
    double  peak_prominence_db, peak_min_db_above_floor, peak_ratio_min;
    float   peak_valid_prom_min_db, peak_valid_prom_max_db;
    /* adaptive tracker quantile of the suppressor's noise-PSD pass (adaptive_q_enable, rain_signal_processor.py:570-576,
     * :634-638, :663-664): q_eff = clip(q - (q - q_min) * rain_ema, q_min, q) with rain_ema the EMA (coefficient alpha) of
     * the "frame excluded from the update" flags, all in float64, cast to float32 where it meets the step */
    int32_t adaptive_q;
    /* moving average over time of the band power before both tracker passes (pre_smooth_frames, :366-379, :690-692;
     * float32 cumulative sum like np.cumsum), <= 1: off, at most APT_MAX_PRE_SMOOTH */
    int32_t pre_smooth_frames;
    double  aq_base, aq_min, aq_alpha;
    /* causal median over time of both passes' noise PSD (median_frames, :381-396, :717-719), <= 1: off, at most APT_MAX_MEDIAN */
    int32_t median_frames;
    /* spectral SNR gating of the oversubtraction (snr_gating_enable, :1050-1077, :433-438):
     * sg[t] = clip(clip(s / (s + snr1), 0, 1) ^ snr_gating_power, 0, 1) (the power only when it is not 1), s = sum(P[mask]) / (sum(N_eff[mask]) + eps), oversub[t] *= 1 - sg[t];
     * mask = bit k of snr_mask for band bin k (union of the mode bands, or the whole band), at most 128 band bins */
    int32_t snr_gating;
    float   snr_gating_snr1;
    float   snr_gating_power;                  /* 1: no power step (:1073-1075) */
    uint32_t snr_mask[4];
    int32_t bypass_classifier;                 /* every frame NOISE, rain_conf 0, noise_conf 1 (:846-857): the suppressor alone */
    /* host pointers, copied at plan creation */
    const double* window;                      /* n_fft analysis window (scipy get_window) */
    const float*  freqs;                       /* n_fft/2+1 bin frequencies as float32 */
} apt_params_t;

/*
 * Output buffers (device pointers, caller-owned; NULL = do not produce).
 * nF = total frames of the batch, K = band_hi-band_lo+1, F = n_fft/2+1, M = n_modes.
 */
typedef struct apt_out_t {
    /* always-required detector outputs */
    int8_t*  frame_class;     /* [nF]  FrameClass NOISE=0 / UNCERTAIN=1 / RAIN=2 */
    float*   rain_conf;       /* [nF] */
    float*   noise_conf;      /* [nF] */
    int32_t* event_idx;       /* [nF]  clip-local frame indices of RAIN frames, ascending, packed at
                                        frame_offsets[c] .. + event_count[c] */
    int32_t* event_count;     /* [n_clips] */
    float*   clip_stats;      /* [n_clips][8]: clip id, rain_frame_count, clip_rain_fraction,
                                 clip_is_rain, clip_rain_conf, median_rain_conf,
                                 mean_noise_floor_db, median_noise_floor_db */
    /* optional planes */
    float*   S;               /* [nF][F][2] complex64 spectrum (state["S"]) */
    float*   P;               /* [nF][F]    power */
    float*   det_noise_psd;   /* [nF][K]    debug["detector_noise_psd"][band] */
    float*   det_noise_lag;   /* [nF][K]    debug["detector_noise_psd_lag"][band] */
    float*   D;               /* [nF][K]    detector input (dB above lagged noise) */
    float*   noise_psd;       /* [nF][K]    state["noise_psd"][band] */
    float*   mode_flux;       /* [M][nF]    raw per-mode flux */
    float*   norm_flux;       /* [M][nF]    primary_mode_flux, support_mode_flux_1.. */
    float*   score;           /* [nF]       mode_flux_score */
    float*   td;              /* [5][nF]    TD features (zero-filled to T) */
    float*   raw;             /* [21][nF]   raw spectral features */
    float*   band_energy;     /* [M+1][nF]  mode-band powers + operating-band energy (float64 sums) */
    uint8_t* gate;            /* [nF]       td_gate_mask */
    float*   x_td;            /* [nS]       zero-phase prefiltered waveform */
    float*   G;               /* [nF][K]    suppressor gain over the band (debug["G"][band]) */
    float*   ratio_med;       /* [nF]       debug["np_ratio_median_t"] */
    float*   S_hat;           /* [nF][F][2] gain-weighted spectrum (state["S_hat"]); needs S */
    float*   peak_ratio;      /* [nF]       det_debug["peak_ratio"]        (peak_features_enable) */
    float*   peak_gate_score; /* [nF]       det_debug["peak_gate_score"]   (all four or none) */
    int32_t* peak_valid_count;/* [nF]       det_debug["peak_valid_count"] */
    int32_t* peak_count_by_mode; /* [M][nF] det_debug["peak_count_by_mode"] */
    float*   y;               /* [nS]       suppressed output audio = ISTFT(S_hat) (state["output_audio"]); needs S_hat */
    float*   td_fast_crest;   /* [nF]       diagnostic: crest factor as the float32 fast path of the TD gate computed it
                                            (the decision uses it only outside the guard band around td_gate_threshold;
                                            frames inside are re-decided by the float64 filter).  Not with td / x_td. */
    float*   snr_mode;        /* [nF]       debug["snr_mode"]: frame SNR over the gating bins (snr_gating_enable; needs G) */
    float*   snr_gate;        /* [nF]       debug["snr_gate"]: clip(snr / (snr + snr1), 0, 1)   (both or none) */
} apt_out_t;

/* which stages a run executes */
#define APT_STAGE_FEATURES 1  /* K1-K3: framing, STFT, power, band energies (BASELINE config 2) */
#define APT_STAGE_FULL     2  /* K1-K9: + TD features, tracker, detection, clip statistics */

int  apt_init(int device_ordinal, apt_ctx** out);
void apt_destroy(apt_ctx* ctx);
const char* apt_last_error(apt_ctx* ctx);
int  apt_abi_version(void);
/* hash of the sources the library was compiled from (set by the build; the Python loader refuses a library whose
   hash differs from the sources beside it) */
const char* apt_source_hash(void);
int  apt_sizeof_params(void);
int  apt_sizeof_out(void);

/* Device self-tests of the branch-free exact float32 primitives the kernels use instead of the IEEE
   division / square root intrinsics: which = 0 square root over every float32 in [1, 2]; which = 1 division on
   n hashed operand pairs.  *mismatches = results that differ from __fsqrt_rn / __fdiv_rn (must be 0). */
int  apt_selftest(apt_ctx* ctx, int which, int64_t n, int64_t* mismatches);

/* Fills *p with the reference's defaults for fs=11162 except the fields that have no default
   (mode bands, window, freqs, SOS): the host must set those.  */
int  apt_params_default(apt_params_t* p);

/* clip_len_samples: host array [n_clips].  The plan owns device scratch sized for the batch. */
int  apt_plan_create(apt_ctx* ctx, const apt_params_t* p, int n_clips,
                     const int64_t* clip_len_samples, apt_plan_t** out);
void apt_plan_destroy(apt_plan_t* plan);
/* host copies of the prefix arrays, each [n_clips+1] */
int  apt_plan_offsets(const apt_plan_t* plan, int64_t* sample_offsets, int64_t* frame_offsets);
int64_t apt_plan_total_frames(const apt_plan_t* plan);
int64_t apt_plan_total_samples(const apt_plan_t* plan);
int64_t apt_plan_scratch_bytes(const apt_plan_t* plan);

/* dev_pcm: concatenated clips on the device (int16 little-endian PCM as on the Mark-3 wire, or the
   float32 waveform the reference's loader hands to processors). */
int  apt_run_i16(apt_plan_t* plan, int stages, const int16_t* dev_pcm, const apt_out_t* out, void* cuda_stream);
int  apt_run_f32(apt_plan_t* plan, int stages, const float* dev_pcm, const apt_out_t* out, void* cuda_stream);

/* number of kernel launches the last apt_run_* call of this plan enqueued */
int  apt_plan_last_launches(const apt_plan_t* plan);

/* Optional per-kernel device timing for benchmarks: when enabled, apt_run_* records CUDA events on
   the launch stream around each kernel group.  After synchronising the stream, apt_plan_kernel_ms
   returns the accumulated milliseconds per group since the last call (and resets them). */
#define APT_KERNEL_STFT 0      /* stft256_kernel */
#define APT_KERNEL_TD 1        /* td_features_kernel */
#define APT_KERNEL_TRK1 2      /* trk1_kernel: tracker pass 1 (serial in time) */
#define APT_KERNEL_FLUX 3      /* flux_kernel: dB normalisation, flux, per-mode sums */
#define APT_KERNEL_BASE 4      /* base_kernel: float64 baselines + normalisation (serial in time) */
#define APT_KERNEL_DECIDE 5    /* decide_kernel + compact_kernel: labels, confidences, event lists */
#define APT_KERNEL_TRK2 6      /* trk2_kernel: tracker pass 2 (serial in time) */
#define APT_KERNEL_DB 7        /* db_kernel: noise-floor dB plane, sums, level-0 histogram */
#define APT_KERNEL_SELECT 8    /* select_hist/scan: median levels 1-2 */
#define APT_KERNEL_FINALIZE 9  /* finalize_kernel */
#define APT_KERNEL_GAIN 10     /* gain_kernel + gain_time_kernel (+ shat_kernel, istft256_kernel): only when G / S_hat / y is requested */
#define APT_N_KERNELS 11
int  apt_plan_enable_timing(apt_plan_t* plan, int enable);
/* Timeline of the pipelined run (benchmarks / profiles): with tracing on, apt_run_* records timing events around
   every launch on its kind's stream.  After synchronising, apt_plan_trace returns for each kernel kind k
   (0 stft, 1 td, 2 trk1, 3 flux, 4 base, 5 decide, 6 trk2, 7 dbsum) and time segment s the start and end of that
   launch in ms since the start of the call: out_ms[(k * 64 + s) * 2 + {0, 1}]; *n_seg = segments of the last run
   (0: it did not run pipelined), *total_ms = duration of the whole call on the caller's stream. */
/* Tensor-core DFT only: 1 if one of its bounded barrier waits gave up (the results of that run are invalid), else 0.
   Synchronises the device. */
int  apt_plan_tc_error(apt_plan_t* plan);
int  apt_plan_enable_trace(apt_plan_t* plan, int enable);
int  apt_plan_trace(apt_plan_t* plan, float* out_ms /* [8][64][2] */, int* n_seg, float* total_ms);
int  apt_plan_kernel_ms(apt_plan_t* plan, float* out_ms /* [APT_N_KERNELS] */);

/* End-to-end convenience path with HOST buffers: copies PCM host->device in clip groups on a copy
   stream while earlier groups compute, runs the full pipeline, and copies frame_class / rain_conf /
   noise_conf / event_idx / event_count / clip_stats back to the given HOST buffers (any may be NULL).
   Synchronises before returning.  Host buffers should be pinned for full PCIe bandwidth. */
int  apt_run_host_i16(apt_plan_t* plan, const int16_t* host_pcm, int8_t* frame_class, float* rain_conf,
                      float* noise_conf, int32_t* event_idx, int32_t* event_count, float* clip_stats);

/* The same end-to-end path for what a processor's run_batch receives: one HOST pointer per clip (pageable memory
   is fine), all int16 (is_f32 = 0) or all float32 (is_f32 = 1; the waveform audio_io.get_input_data hands out,
   audio_io.py:421-432), clip c holding the plan's clip_len_samples[c] samples.  Helper threads stage the clips
   of each group into a small ring of pinned buffers owned by the plan while the previous group is on the bus and
   earlier groups compute; results are copied back as in apt_run_host_i16.  This is the entry point behind
   RainDetectorProcessor.run_batch / NoiseProcessor.run_batch (the replacement of the per-file proc.run loop,
   audio_processing_framework.py:183-207).  Synchronises before returning.
   Environment: APT_HOST_GROUPS (clip groups, default 24), APT_STAGE_THREADS (staging threads, default min(16, cores/2)). */
int  apt_run_host_clips(apt_plan_t* plan, const void* const* clip_ptrs, int is_f32, int8_t* frame_class, float* rain_conf,
                        float* noise_conf, int32_t* event_idx, int32_t* event_count, float* clip_stats);

/* ---------------------------------------------------------------------------------------------
 * Drop-size-distribution emulator: replaces DsdProcessingEmualtor.process_audio_data
 * (host_analysis/device_dsd_processing_emulator.py:257-314), the compute path of
 * transform.process_audio_file_dsd (transform.py:251-313), for a batch of clips.
 * ------------------------------------------------------------------------------------------- */
#define APT_DSD_OUT 100   /* 32 drop-size bins + 30 peak-frequency slots + 38 FFT energies per minute */
typedef struct apt_dsd_params_t {
    int32_t fs, frame_length, hop_length, apply_window;   /* constructor arguments (:16-31) */
    const double* window;                                  /* frame_length values when apply_window (host) */
} apt_dsd_params_t;
/* clip_len / ts: host arrays [n_clips] (ts = start time of each clip in seconds, as process_audio_data's ts).
   dev_pcm: concatenated int16 clips on the device (scaled /32768 as parse.pcm_to_float, :670).
   dev_out: [n_clips][max_minutes][100] float64, dev_n_minutes: [n_clips] int32 (rows actually produced).
   max_minutes must be >= ceil(len / (fs * 60)) of the longest clip.  Runs on the given stream; synchronises it
   before returning (per-frame scratch is released). */
int  apt_dsd_run_i16(apt_ctx* ctx, const apt_dsd_params_t* p, int n_clips, const int64_t* clip_len, const double* ts,
                     const int16_t* dev_pcm, double* dev_out, int32_t* dev_n_minutes, int max_minutes, void* cuda_stream);

/* ---------------------------------------------------------------------------------------------
 * Band noise estimator: replaces the per-frame loop of BandNoiseEstimatorProcessor.run
 * (edge/band_noise_processor.py:82-281) over BandNoiseEstimator.process_frame
 * (edge/band_noise_estimator.py:770-986) for a batch of clips; float64 configuration, hop == frame_len.
 * The host resolves the configuration (filter design, bins, ratios) exactly as the reference's constructors do.
 * ------------------------------------------------------------------------------------------- */
#define APT_BNE_MAX_SOS 8
#define APT_BNE_MAX_S 8
#define APT_BNE_MAX_BANDS 8
#define APT_BNE_FRAME_F 12   /* M_band, E_band, N_E, N_E_raw, G_mag, M_clean, noise_effective_q, M_band_fft, E_band_fft,
                                E_hpf, N_sub, fft_rain_frame */
#define APT_BNE_STATS 16     /* noise/rain/total energy sums, noise/rain/total frame counts, buffer valid / min valid /
                                underflow counts, frames_since_noise_update, learned, replenished, noise_effective_q */
typedef struct apt_bne_params_t {
    int32_t fs, N, sub_len, S;
    int32_t ns_h, ns_b, warm, pad0;            /* sections of the HPF / BPF, warm-up samples of a filter segment */
    double  sos_h[APT_BNE_MAX_SOS][6], zi_h[APT_BNE_MAX_SOS][2];   /* scipy butter(...,"sos"), sosfilt_zi */
    double  sos_b[APT_BNE_MAX_SOS][6], zi_b[APT_BNE_MAX_SOS][2];
    int32_t n_bands, band_b0[APT_BNE_MAX_BANDS], band_b1[APT_BNE_MAX_BANDS], prim_b0, prim_b1, mask_b0, mask_b1, pad1;
    double  M_ratio, N_ratio, D_ratio, band_rise_db, excess_rise_db, min_Ehpf, min_Eband, dE_thr;
    int32_t k_subframes, use_dE, use_D, pad2;
    int32_t W, W_min, ttl, smooth, learn_all, replenish, replenish_only_not_full, q_adapt;
    double  q, ema_alpha, beta, gain_floor, eps, att_dry, att_wet, release, repl_q, q_repl_alpha, q_norm_alpha;
} apt_bne_params_t;
int  apt_sizeof_bne_params(void);
/* dev_pcm: concatenated clips (int16 scaled /32767 as audio_io.safe_to_float when is_f32 == 0, else float32).
   dev_frame_out [nF][APT_BNE_FRAME_F] f64, dev_mask [nF] u8 (bit s = rain_submask[s]), dev_subE [nF][APT_BNE_MAX_S] f64,
   dev_stats [n_clips][APT_BNE_STATS] f64; nF = sum over clips of len / N.  Synchronises the stream before returning. */
int  apt_bne_run(apt_ctx* ctx, const apt_bne_params_t* p, int n_clips, const int64_t* clip_len, const void* dev_pcm, int is_f32,
                 double* dev_frame_out, uint8_t* dev_mask, double* dev_subE, double* dev_stats, void* cuda_stream);

/* ---------------------------------------------------------------------------------------------
 * Legacy "RoE" rain detector.  Replaces, for a batch of clips, the compute of
 * `rain_detection_algo` (edge/dsp_rain_detection.py:2566-2575): `analyse_raw_audio` per 2-second part
 * (:2230-2562, :2603-2636) with `calculate_pulse_characteristics` (:657-767), `compute_novelty_spectrum_new`
 * (:1924-1955), `find_peaks_in_frequency_range` (:1649-1698), and the clip-level FP / FN combination
 * (:2638-2731).  The host resolves `configure_parameters` (:1298-1391) into apt_roe_params_t (filter design by
 * scipy.signal.butter stays on the host) and lists the parts; `max_harmonics` -- module state of the reference
 * that survives between calls (:1141, :1394-1403) -- goes in and comes out explicitly.
 * ------------------------------------------------------------------------------------------- */
#define APT_ROE_FRAME_F 8   /* per frame slot: raining, kurtosis, crest_factor, diff_energy, energy, min_energy, Nov0, novt */
#define APT_ROE_PART_F 5    /* per part: frain_mean, natural-range flag, max_harmonics set by the part (0: none), drops, td peaks */
#define APT_ROE_CLIP_F 5    /* per clip: returned rain_drops, rain_drop_count, rain_peaks_count, rain_drop_count_mod, raining */
typedef struct apt_roe_params_t {
    int32_t n_fft, hop;                 /* 256 / 128 */
    int32_t M, wl;                      /* local-average half window (min_average_len), values averaged = max(3, M / 6) */
    int32_t max_peaks, want_td;         /* want_td = handle_fp || handle_fn */
    int32_t ns_in, ns_td;               /* biquad sections of the two band-pass filters */
    double  sos_in[8][6];               /* butter(8, op_freq_range, "bandpass") */
    double  sos_td[4][6];               /* butter(4, [400, 900], "band") */
    double  window[256];                /* scipy get_window("hann", 256) */
    double  fs;                         /* 11162: analyse_raw_audio's own default, whatever sample_rate says (:2236) */
    double  f_natural, op_lo, op_hi, nat_lo, nat_hi, search0_lo, search0_hi;
    double  rain_thr[6], rain_thr_hn;
    double  kurtosis_thr, crest_thr, diff_energy_thr;
    int32_t handle_fp, handle_fn, rain_drop_threshold, rain_drop_max_thr, rain_peaks_min_thr, rain_peaks_max_thr;
} apt_roe_params_t;
int  apt_sizeof_roe_params(void);
/* Parts are listed clip by clip in time order: part_clip ascending, part_start = first sample (absolute index into
   dev_pcm), part_len >= fs samples and <= 254 * hop.  A part has T = 1 + part_len / hop frames and T + 1 frame slots
   (the reference appends one zero per part); slots of all parts are concatenated.
   dev_frame_out [sum(T + 1)][APT_ROE_FRAME_F] f64, dev_part_out [n_parts][APT_ROE_PART_F] f64,
   dev_clip_out [n_clips][APT_ROE_CLIP_F] f64 (a clip without parts gets zeros).  Synchronises the stream. */
int  apt_roe_run(apt_ctx* ctx, const apt_roe_params_t* p, int n_clips, const void* dev_pcm, int is_f32, int n_parts,
                 const int32_t* part_clip, const int64_t* part_start, const int32_t* part_len, int max_harmonics_in,
                 double* dev_frame_out, double* dev_part_out, double* dev_clip_out, int* max_harmonics_out, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* APT_B200_H */
