"""Configuration of the spectral noise / rain-frame engine and its resolution to the C ABI.

Mirrors the reference's configuration surface:
  * ``NoiseProcessorConfig``  -- edge/rain_signal_processor.py:19-188 (same field names/defaults)
  * ``build_noise_config``    -- edge/rain_signal_processor.py:202-255 (flat > suppressor > defaults,
                                 detector kept nested, legacy fmin/fmax -> operating_band)
  * detector lookups          -- edge/rain_frame_classifier.py:135-148 (detector dict > cfg attr > default)
``resolve_params`` turns that configuration into the POD ``apt_params_t`` (include/apt_b200.h),
applying the casts numpy applies when the reference mixes Python floats with float32 arrays.
Flag combinations the CUDA path does not implement are rejected loudly -- never emulated on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field, fields
from typing import Any, Dict, Tuple

import numpy as np
import scipy.signal as spsig

from . import _lib


@dataclass
class NoiseProcessorConfig:
    # core STFT geometry
    fs: int = 11162
    n_fft: int = 256
    hop: int = 128
    # pre-filter
    hp_cutoff_hz: float = 350.0
    hp_order: int = 4
    pre_filter_mode: str = "highpass"
    bp_order: int = 4
    operating_band: Tuple[float, float] = (400.0, 3500.0)
    # noise tracking
    q: float = 0.25
    win_sec: float = 0.5
    adaptive_q_enable: bool = False
    adaptive_q_min: float = 0.10
    adaptive_q_alpha: float = 0.95
    median_frames: int = 0
    eps: float = 1e-9
    noise_psd_max_ratio: float = 1.0
    use_lagged_noise_psd: bool = False
    # suppressor gain
    oversub_base: float = 1.0
    oversub_max: float = 3.0
    gain_floor: float = 0.0
    gain_ceil: float = 1.0
    gain_mode: str = "sqrt_sub"
    gain_smooth_alpha: float = 0.7
    adaptive_gain_enable: bool = True
    gain_freq_smooth_enable: bool = True
    gain_freq_kernel: Tuple[float, ...] = (0.2, 0.6, 0.2)
    pre_smooth_frames: int = 0
    ema_up: float = 0.6
    ema_down: float = 0.95
    snr_gating_enable: bool = False
    snr_gating_snr1: float = 1.0
    snr_gating_power: float = 1.0
    snr_gating_use_mode_bands: bool = True
    detector_use_noise_norm: bool = True
    detector_noise_norm_mode: str = "log_sub"
    suppressor_bypass: bool = False
    classifier_only_mode: bool = False
    debug_enable: bool = False
    debug_frame_decim: int = 1
    dump_features: bool = False
    feature_decim: int = 1
    process_dtype: str = "float32"
    compute_output_audio: bool = False
    return_filtered_audio: bool = False
    return_debug: bool = False
    return_detector_debug: bool = False
    return_spectra: bool = False
    return_noise_psd: bool = False
    suppressor: Dict[str, Any] = field(default_factory=dict)
    detector: Dict[str, Any] = field(default_factory=dict)


def build_noise_config(sample_rate: int, params: Dict[str, Any]) -> NoiseProcessorConfig:
    cfg = NoiseProcessorConfig(fs=int(sample_rate))
    names = {f.name for f in fields(NoiseProcessorConfig)}
    merged = dict(params)
    nested_sup = merged.get("suppressor")
    if isinstance(nested_sup, dict):
        cfg.suppressor = dict(nested_sup)
        merged = {**nested_sup, **merged}
    nested_det = merged.get("detector")
    if isinstance(nested_det, dict):
        cfg.detector = dict(nested_det)
    if "operating_band" not in merged:
        lo, hi = merged.get("fmin"), merged.get("fmax")
        if lo is not None and hi is not None:
            merged["operating_band"] = (float(lo), float(hi))
    for key, value in merged.items():
        if key not in names:
            continue
        if key == "operating_band" and isinstance(value, (list, tuple)) and len(value) == 2:
            value = (float(value[0]), float(value[1]))
        elif key == "gain_freq_kernel":
            value = tuple(float(v) for v in value)
        setattr(cfg, key, value)
    cfg.operating_band = (float(cfg.operating_band[0]), float(cfg.operating_band[1]))
    return cfg


class DetectorView:
    """detector dict > cfg attribute > default."""

    def __init__(self, cfg: NoiseProcessorConfig):
        self.cfg = cfg
        self.det = cfg.detector if isinstance(cfg.detector, dict) else {}

    def get(self, name, default=None):
        if name in self.det:
            return self.det[name]
        if hasattr(self.cfg, name):
            return getattr(self.cfg, name)
        return default

    def has(self, name):
        return name in self.det or hasattr(self.cfg, name)


def validate_config(cfg: NoiseProcessorConfig) -> None:
    """Same checks (and exception types) as _validate_rain_cfg / _validate_suppressor_cfg
    (rain_frame_classifier.py:165-176, rain_signal_processor.py:301-333)."""
    if not DetectorView(cfg).has("mode_bands"):
        raise AttributeError("RainFrameClassifierMixin missing required detector fields: ['mode_bands']. "
                             "Provide them under cfg.detector (preferred) or as flat cfg attributes.")
    lo, hi = cfg.operating_band
    if not (np.isfinite(lo) and np.isfinite(hi) and 0.0 < float(lo) < float(hi)):
        raise ValueError(f"Invalid operating_band: {cfg.operating_band!r}")
    if int(cfg.n_fft) <= 0 or int(cfg.hop) <= 0:
        raise ValueError(f"Invalid STFT params n_fft={cfg.n_fft}, hop={cfg.hop}")
    if int(cfg.hop) > int(cfg.n_fft):
        raise ValueError(f"hop ({cfg.hop}) should not exceed n_fft ({cfg.n_fft})")
    if not (0.0 <= float(cfg.gain_floor) <= float(cfg.gain_ceil) <= 1.0):
        raise ValueError(f"Invalid gain bounds: floor={cfg.gain_floor}, ceil={cfg.gain_ceil}")
    if float(cfg.oversub_base) <= 0.0 or float(cfg.oversub_max) <= 0.0:
        raise ValueError(f"Invalid oversub params: base={cfg.oversub_base}, max={cfg.oversub_max}")
    if float(cfg.oversub_max) < float(cfg.oversub_base):
        raise ValueError(f"oversub_max ({cfg.oversub_max}) must be >= oversub_base ({cfg.oversub_base})")
    if not (0.0 <= float(cfg.gain_smooth_alpha) <= 1.0):
        raise ValueError(f"Invalid gain_smooth_alpha: {cfg.gain_smooth_alpha}")


def _contiguous_bins(mask):
    idx = np.flatnonzero(mask)
    if idx.size == 0:
        return 1, 0
    if idx[-1] - idx[0] + 1 != idx.size:
        raise NotImplementedError("non-contiguous bin mask")
    return int(idx[0]), int(idx[-1])


_UNSUPPORTED_DETECTOR_FLAGS = ("feature_dump_include_peak_payload", "flux_modes_winsor_enable",
                               "td_envelope_features_enable", "clip_spectral_occupancy_enable")


class ResolvedParams:
    """apt_params_t plus the host-side tables it points to (kept alive with it)."""

    def __init__(self, cfg: NoiseProcessorConfig, sample_rate: int, clip_rain_min_frames: int = 1,
                 fft_f64: bool = True):
        validate_config(cfg)
        dv = DetectorView(cfg)
        f32 = np.float32
        for flag in _UNSUPPORTED_DETECTOR_FLAGS:
            if bool(dv.get(flag, False)):
                raise NotImplementedError(f"detector.{flag}=True is not implemented on the CUDA path")
        if str(cfg.process_dtype).lower() != "float32":
            raise NotImplementedError("process_dtype='float64' is not implemented on the CUDA path")
        if int(cfg.pre_smooth_frames or 0) > _lib.MAX_PRE_SMOOTH or int(cfg.median_frames or 0) > _lib.MAX_MEDIAN - 1:
            raise NotImplementedError(f"pre_smooth_frames > {_lib.MAX_PRE_SMOOTH} / median_frames > {_lib.MAX_MEDIAN - 1} "
                                      "are not implemented on the CUDA path")
        if str(dv.get("td_input_mode", "default")).lower() != "default":
            raise NotImplementedError("td_input_mode other than 'default' is not implemented on the CUDA path")

        P = _lib.AptParams()
        P.abi_version = _lib.ABI_VERSION
        sr = int(sample_rate)
        n_fft, hop = int(cfg.n_fft), int(cfg.hop)
        P.fs, P.n_fft, P.hop = sr, n_fft, hop
        self.freqs = np.ascontiguousarray(np.asarray(np.fft.rfftfreq(n=n_fft, d=1.0 / sr), dtype=np.float32))
        self.window = np.ascontiguousarray(spsig.get_window("hann", n_fft, fftbins=True).astype(np.float64))
        op_lo, op_hi = cfg.operating_band
        self.band_mask = (self.freqs >= op_lo) & (self.freqs <= op_hi)
        if not np.any(self.band_mask):
            raise ValueError(f"operating_band {cfg.operating_band} does not overlap the provided frequency grid")
        P.band_lo, P.band_hi = _contiguous_bins(self.band_mask)

        mode_bands = dv.get("mode_bands")
        if mode_bands is None:
            raise AttributeError("Missing required detector param: mode_bands")
        mode_bands = tuple((float(a), float(b)) for (a, b) in mode_bands)
        if len(mode_bands) < 4:
            raise ValueError("Fixed-band rain decision requires at least 4 mode bands: "
                             "mode 0 as primary and modes 1, 2, 3 as support")
        if len(mode_bands) > _lib.MAX_MODES:
            raise NotImplementedError(f"more than {_lib.MAX_MODES} mode bands")
        weights = dv.get("mode_weights")
        if weights is not None:
            weights = tuple(float(w) for w in weights)
            if len(weights) != len(mode_bands):
                raise ValueError(f"mode_weights length ({len(weights)}) must match mode_bands length ({len(mode_bands)})")
        self.mode_bands = mode_bands
        P.n_modes = len(mode_bands)
        in_band = self.freqs[self.band_mask]
        wide = self.freqs.astype(np.float64)
        any_mode = False
        for i in range(_lib.MAX_MODES):
            P.mode_lo[i], P.mode_hi[i], P.mode_band_lo[i], P.mode_band_hi[i], P.mode_weight[i] = 1, 0, 1, 0, 1.0
        for i, (lo, hi) in enumerate(mode_bands):
            P.mode_lo[i], P.mode_hi[i] = _contiguous_bins((wide >= lo) & (wide <= hi))
            P.mode_band_lo[i], P.mode_band_hi[i] = _contiguous_bins((in_band >= lo) & (in_band <= hi))
            any_mode = any_mode or P.mode_band_lo[i] <= P.mode_band_hi[i]
            if weights is not None:
                P.mode_weight[i] = weights[i]
        if P.mode_band_lo[0] > P.mode_band_hi[0]:
            raise ValueError(f"primary mode band {mode_bands[0]} has no bins inside operating_band {cfg.operating_band}")
        if not any_mode:
            raise ValueError("No mode band overlaps the operating band")

        # --- noise-PSD tracker constants
        frames_per_sec = float(sr) / float(hop)
        W = max(10, int(cfg.win_sec * frames_per_sec))
        eta = float(np.clip(float(2.0 / max(W + 1, 2)), 1e-4, 1.0))
        alpha = float(cfg.ema_down)
        maxr = float(cfg.noise_psd_max_ratio)
        maxr = 1.0 if not np.isfinite(maxr) else float(np.clip(maxr, 0.0, 1.0))
        q = float(cfg.q)
        P.trk_eta, P.trk_scale_alpha, P.trk_one_minus_alpha = f32(eta), f32(alpha), f32(1.0 - alpha)
        P.trk_step_floor = f32(float(max(cfg.eps, 1e-9)))
        P.trk_q, P.trk_neg_one_minus_q, P.trk_maxr = f32(q), f32(-(1.0 - q)), f32(maxr)
        # adaptive quantile of pass 2 (rain_signal_processor.py:570-576); pass 1 excludes no frame, so its q never moves
        P.adaptive_q = int(bool(cfg.adaptive_q_enable))
        P.aq_base = q
        P.aq_min = float(np.clip(float(cfg.adaptive_q_min), 1e-4, q))
        P.aq_alpha = float(np.clip(float(cfg.adaptive_q_alpha), 0.0, 1.0))
        P.pre_smooth_frames = int(cfg.pre_smooth_frames or 0)      # rain_signal_processor.py:690-692
        P.median_frames = int(cfg.median_frames or 0)              # :717-719
        P.ema_up, P.ema_down = float(cfg.ema_up), float(cfg.ema_down)
        P.warmup_need = max(10, W // 2)
        P.eps_f32 = f32(cfg.eps)
        P.detector_use_noise_norm = int(bool(dv.get("detector_use_noise_norm", True)))
        P.norm_ratio_db = int(str(cfg.detector_noise_norm_mode).lower() == "ratio_db")

        # --- flux baseline constants (Python doubles)
        eps = float(dv.get("eps", 1e-9))
        q_pct = float(np.clip(float(dv.get("mode_flux_norm_q", 20.0)), 0.0, 100.0))
        norm_min = max(float(dv.get("mode_flux_norm_min", 1.0)), eps)
        det_fs = float(dv.get("sample_rate", dv.get("fs", 11162)))
        det_fps = float(max(det_fs / max(float(dv.get("hop", 128)), 1.0), 1e-6))
        Wb = max(3, int(round(float(dv.get("mode_flux_norm_win_sec", 0.5)) * det_fps)))
        b_eta = float(np.clip(2.0 / max(Wb + 1, 2), 1e-4, 1.0))
        P.bl_q = float(np.clip(q_pct, 0.0, 100.0)) / 100.0
        P.bl_eta = b_eta
        P.bl_scale_alpha = float(np.clip(1.0 - b_eta, 0.0, 0.9999))
        P.bl_floor = float(max(norm_min, 1e-12))
        P.norm_enable = int(bool(dv.get("mode_flux_norm_enable", True)))
        P.norm_min_f32 = f32(norm_min)

        # --- decision
        legacy = float(dv.get("new_rain_mode12_flux_min", 2.6))
        P.thr_primary = f32(float(dv.get("new_rain_primary_flux_min", 1.8)))
        P.thr_m1 = f32(float(dv.get("new_rain_mode1_flux_min", legacy)))
        P.thr_m2 = f32(float(dv.get("new_rain_mode2_flux_min", legacy)))
        P.thr_m3 = f32(float(dv.get("new_rain_mode3_flux_min", 3.0)))
        P.min_support = int(dv.get("new_rain_min_support_count", 2))
        P.td_gate_thr = f32(float(dv.get("td_gate_threshold", 2.5)))
        upper = dv.get("td_kurtosis_upper_threshold", None)
        P.has_kurt_upper = int(upper is not None)
        P.kurt_upper = f32(float(upper)) if upper is not None else f32(0.0)
        P.noise_hi = f32(float(dv.get("noise_hi", 0.80)))
        P.mode_flux_noise_max = f32(max(float(dv.get("mode_flux_noise_max", 1.5)), 0.0))

        # --- zero-phase TD prefilter
        td_mode = str(dv.get("td_prefilter_mode", dv.get("pre_filter_mode", "none"))).lower()
        self.sos = None
        if bool(dv.get("td_apply_input_prefilter", True)) and td_mode not in ("", "none"):
            self.sos = prefilter_sos(cfg, sr, td_mode)
        if self.sos is None:
            P.n_sos, P.padlen = 0, 0
        else:
            ns = int(self.sos.shape[0])
            if ns > _lib.MAX_SOS:
                raise NotImplementedError(f"prefilter with {ns} second-order sections")
            ntaps = 2 * ns + 1
            ntaps -= min(int((self.sos[:, 2] == 0).sum()), int((self.sos[:, 5] == 0).sum()))
            P.n_sos, P.padlen = ns, 3 * ntaps
            zi = spsig.sosfilt_zi(self.sos)
            for s in range(ns):
                for j in range(6):
                    P.sos[s][j] = float(self.sos[s, j])
                P.zi[s][0], P.zi[s][1] = float(zi[s, 0]), float(zi[s, 1])
        P.eps_f64 = eps

        # --- TD block-energy features
        P.blk_len = int(max(1, int(dv.get("td_block_energy_len", 8))))
        bhop = dv.get("td_block_energy_hop", None)
        P.blk_hop = max(1, int(bhop)) if bhop is not None else P.blk_len
        if P.blk_hop != P.blk_len:
            raise NotImplementedError("td_block_energy_hop != td_block_energy_len is not implemented on the CUDA path")
        P.blk_post_pre = int(dv.get("td_block_energy_post_pre_blocks", 4))
        P.blk_smooth = int(bool(dv.get("td_block_energy_smooth_enable", True)))

        # --- raw spectral features
        if not bool(dv.get("raw_spectral_shape_enable", True)):
            self.raw_enabled = False
        else:
            self.raw_enabled = True
        low = dv.get("raw_spectral_low_band", (50.0, 200.0))
        rain = dv.get("raw_spectral_rain_band", (400.0, 800.0))
        P.low_lo, P.low_hi = _contiguous_bins((wide >= max(float(low[0]), eps)) & (wide < float(low[1])))
        P.rain_lo, P.rain_hi = _contiguous_bins((wide >= float(rain[0])) & (wide <= float(rain[1])))
        P.rolloff_fraction = float(dv.get("raw_spectral_rolloff_fraction", 0.85))
        P.suppressor_bypass = int(bool(cfg.suppressor_bypass))
        P.clip_rain_min_frames = int(max(1, int(clip_rain_min_frames)))
        # True / 1: float64 FFT (reference arithmetic); False / 0: float32 FFT; "tc" / 2: tensor-core DFT (tolerance path)
        P.fft_f64 = 2 if fft_f64 in ("tc", 2) else int(bool(fft_f64))

        # --- suppressor gain (_compute_gain, rain_signal_processor.py:400-533).  noise_conf is binary on this
        # path (1 - rain_conf), so the per-frame scalars take two values; they are formed here with the same
        # numpy float32 / Python-float promotions the reference applies.
        P.bypass_classifier = int(bool(dv.get("bypass_classifier", False)))   # rain_signal_processor.py:846-857
        P.snr_gating = int(bool(cfg.snr_gating_enable))           # rain_signal_processor.py:1050-1077
        if P.snr_gating:
            pwr = float(cfg.snr_gating_power)
            P.snr_gating_power = f32(pwr) if (pwr != 1.0 and np.isfinite(pwr) and pwr > 0.0) else f32(1.0)   # :1073-1075
            Kb = int(P.band_hi - P.band_lo + 1)
            if Kb > 128:
                raise NotImplementedError("snr_gating_enable needs an operating band of at most 128 bins")
            mask = np.zeros(Kb, dtype=bool)
            if bool(cfg.snr_gating_use_mode_bands):
                for i in range(int(P.n_modes)):
                    if P.mode_band_lo[i] <= P.mode_band_hi[i]:
                        mask[P.mode_band_lo[i]:P.mode_band_hi[i] + 1] = True
            if not mask.any():
                mask[:] = True
            for k in np.flatnonzero(mask):
                P.snr_mask[int(k) >> 5] |= (1 << (int(k) & 31))
            P.snr_gating_snr1 = f32(max(1e-9, float(cfg.snr_gating_snr1)))
        mode = str(cfg.gain_mode).lower()
        P.gain_mode = 1 if mode == "wiener" else 0
        P.adaptive_gain = int(bool(cfg.adaptive_gain_enable))
        P.gain_freq_smooth = int(bool(cfg.gain_freq_smooth_enable))
        P.use_lagged_noise_psd = int(bool(cfg.use_lagged_noise_psd))
        th = 0.7
        denom = max(1e-9, 1.0 - th)
        nc = np.array([1.0, 0.0], dtype=np.float32)
        if P.adaptive_gain:
            eff = np.clip((nc - th) / denom, 0.0, 1.0)
            oversub = cfg.oversub_base + eff * (cfg.oversub_max - cfg.oversub_base)
        else:
            eff = np.zeros(2, dtype=np.float32)
            oversub = np.full(2, float(cfg.oversub_base), dtype=np.float32)
        P.oversub_noise, P.oversub_rain = f32(oversub[0]), f32(oversub[1])
        P.gain_floor, P.gain_ceil = f32(cfg.gain_floor), f32(cfg.gain_ceil)
        kernel = np.asarray(cfg.gain_freq_kernel, dtype=np.float32).reshape(-1)
        if kernel.size < 1:
            kernel = np.array([1.0], dtype=np.float32)
        kernel = kernel / (kernel.sum() + 1e-12)
        if kernel.size > _lib.MAX_GAIN_TAPS or kernel.size % 2 == 0:
            raise NotImplementedError("gain_freq_kernel must have an odd number of taps, at most %d" % _lib.MAX_GAIN_TAPS)
        P.n_gain_taps = int(kernel.size)
        for i in range(_lib.MAX_GAIN_TAPS):
            P.gain_taps[i] = f32(kernel[i]) if i < kernel.size else f32(0.0)
        alpha_base = float(np.clip(cfg.gain_smooth_alpha, 0.0, 1.0))
        eff_nc = (np.float32(1.0) - th) / denom                  # np.float32, as for noise_conf[t] = 1
        a_noise = alpha_base * eff_nc
        P.alpha_noise, P.one_minus_alpha_noise = f32(a_noise), f32(1.0 - a_noise)
        P.alpha_base, P.one_minus_alpha_base = f32(alpha_base), f32(1.0 - alpha_base)
        P.gain_eps_f32 = f32(cfg.eps)

        # --- optional peak-structure features (rain_frame_classifier.py:670-683)
        self.peak_features = bool(dv.get("peak_features_enable", False))
        P.peak_top_p = max(1, int(dv.get("peak_top_p", 6)))
        P.primary_top_m = max(1, int(dv.get("primary_top_m", 3)))
        P.peak_prominence_db = float(dv.get("peak_prominence_db", 3.0))
        P.peak_min_db_above_floor = float(dv.get("peak_min_db_above_floor", 6.0))
        P.peak_ratio_min = float(np.clip(float(dv.get("peak_ratio_min", 0.50)), 0.0, 1.0))
        vmin = float(dv.get("peak_valid_prom_min_db", 3.0))
        vmax = max(vmin, float(dv.get("peak_valid_prom_max_db", 6.0)))
        P.peak_valid_prom_min_db, P.peak_valid_prom_max_db = f32(vmin), f32(vmax)
        P.window = self.window.ctypes.data_as(C.c_void_p)
        P.freqs = self.freqs.ctypes.data_as(C.c_void_p)
        self.c = P
        self.cfg = cfg
        self.sample_rate = sr
        self.K = P.band_hi - P.band_lo + 1
        self.F = n_fft // 2 + 1
        self.M = P.n_modes


def prefilter_sos(cfg: NoiseProcessorConfig, sr: int, mode: str):
    """edge/rain_signal_processor.py:347-364 (scipy.signal.butter, SOS form)."""
    nyq = 0.5 * sr
    if mode == "bandpass":
        lo = np.clip(float(cfg.operating_band[0]), 1e-3, nyq * 0.999)
        hi = np.clip(float(cfg.operating_band[1]), lo + 1e-3, nyq * 0.999)
        return spsig.butter(int(getattr(cfg, "bp_order", cfg.hp_order)), [lo / nyq, hi / nyq],
                            btype="bandpass", output="sos")
    if mode == "highpass" and cfg.hp_cutoff_hz > 0:
        return spsig.butter(cfg.hp_order, np.clip(cfg.hp_cutoff_hz / nyq, 1e-4, 0.9999),
                            btype="highpass", output="sos")
    return None
