"""Python front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package never does.

It resolves a reference-style ``params`` dict the way the reference does
(edge/rain_signal_processor.py:202-255 ``build_noise_config`` and
edge/rain_frame_classifier.py:135-148 ``_dget``), packs the constants into the C
struct of oracle/apt_oracle.c with the casts numpy (NEP 50) would apply, and returns
the same results/state dictionaries as ``RainDetectorProcessor.run``
(edge/rain_signal_processor.py:1223-1344).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.signal as spsig

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libapt_oracle.so")
MAXM, MAXS = 8, 8

RAW_NAMES = (
    "raw_spectral_centroid_hz", "raw_spectral_bandwidth_hz", "raw_low_freq_ratio",
    "raw_rain_band_ratio", "raw_mode_band_ratio_0", "raw_mode_band_ratio_1",
    "raw_mode_band_ratio_2", "raw_mode_band_ratio_3", "raw_mode_band_ratio_4",
    "raw_mode_band_entropy", "raw_mode_band_std", "raw_mode_band_max_ratio",
    "raw_spectral_flatness", "raw_spectral_rolloff_hz", "raw_dominant_freq_hz",
    "raw_frame_energy", "raw_cepstrum_coeff_0", "raw_cepstrum_coeff_1",
    "raw_cepstrum_coeff_2", "raw_cepstrum_coeff_3", "raw_cepstrum_coeff_4")
TD_NAMES = ("td_crest_factor", "td_kurtosis", "td_block_energy_crest",
            "td_block_peak_width_50", "td_block_post_pre_energy_ratio")


class OrcParams(C.Structure):
    _fields_ = [
        ("fs", C.c_int32), ("n_fft", C.c_int32), ("hop", C.c_int32),
        ("band_lo", C.c_int32), ("band_hi", C.c_int32), ("n_modes", C.c_int32),
        ("mode_lo", C.c_int32 * MAXM), ("mode_hi", C.c_int32 * MAXM),
        ("mode_in_band_lo", C.c_int32 * MAXM), ("mode_in_band_hi", C.c_int32 * MAXM),
        ("mode_weight", C.c_double * MAXM),
        ("trk_eta", C.c_float), ("trk_scale_alpha", C.c_float), ("trk_one_minus_alpha", C.c_float),
        ("trk_step_floor", C.c_float), ("trk_q", C.c_float), ("trk_neg_one_minus_q", C.c_float),
        ("trk_maxr", C.c_float),
        ("ema_up", C.c_double), ("ema_down", C.c_double),
        ("warmup_need", C.c_int32), ("eps_f32", C.c_float),
        ("detector_use_noise_norm", C.c_int32), ("norm_ratio_db", C.c_int32),
        ("bl_q", C.c_double), ("bl_eta", C.c_double), ("bl_scale_alpha", C.c_double), ("bl_floor", C.c_double),
        ("norm_enable", C.c_int32), ("norm_min_f32", C.c_float),
        ("thr_primary", C.c_float), ("thr_m1", C.c_float), ("thr_m2", C.c_float), ("thr_m3", C.c_float),
        ("min_support", C.c_int32), ("td_gate_thr", C.c_float), ("has_kurt_upper", C.c_int32),
        ("kurt_upper", C.c_float), ("noise_hi", C.c_float), ("mode_flux_noise_max", C.c_float),
        ("n_sos", C.c_int32), ("padlen", C.c_int32),
        ("sos", (C.c_double * 6) * MAXS), ("zi", (C.c_double * 2) * MAXS),
        ("eps_f64", C.c_double),
        ("blk_len", C.c_int32), ("blk_hop", C.c_int32), ("blk_post_pre", C.c_int32), ("blk_smooth", C.c_int32),
        ("low_lo", C.c_int32), ("low_hi", C.c_int32), ("rain_lo", C.c_int32), ("rain_hi", C.c_int32),
        ("rolloff_fraction", C.c_double),
        ("suppressor_bypass", C.c_int32), ("adaptive_q", C.c_int32),
        ("aq_base", C.c_double), ("aq_min", C.c_double), ("aq_alpha", C.c_double),
        ("pre_smooth_frames", C.c_int32), ("median_frames", C.c_int32),
        ("bypass_classifier", C.c_int32), ("reserved", C.c_int32),
    ]


class OrcOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "S", "P_band", "N1_band", "Nlag_band", "D_band", "N2_band", "mode_flux", "flux_modes",
        "baseline", "norm_flux", "score", "x_td", "td", "raw", "gate", "frame_class",
        "rain_conf", "noise_conf")]


_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
            os.path.join(_HERE, "apt_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        assert _lib.orc_sizeof_params() == C.sizeof(OrcParams), "oracle struct mismatch"
        _lib.orc_num_frames.restype = C.c_int64
        _lib.orc_num_frames.argtypes = [C.c_int64, C.c_int]
        _lib.orc_np_sum_f32.restype = C.c_float
        _lib.orc_np_sum_f32.argtypes = [C.c_void_p, C.c_int64]
        _lib.orc_baseline.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_double,
                                      C.c_double, C.c_void_p]
        _lib.orc_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        _lib.orc_process_batch_i16.argtypes = [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int]
        _lib.orc_noise_db.argtypes = [C.c_void_p, C.c_int64, C.c_float, C.c_void_p]
        _lib.orc_hann_periodic.argtypes = [C.c_void_p, C.c_int]
    return _lib


# ----------------------------------------------------------------------------
# parameter resolution (follows the reference, not the product package)
# ----------------------------------------------------------------------------
_CFG_DEFAULTS = dict(
    n_fft=256, hop=128, hp_cutoff_hz=350.0, hp_order=4, pre_filter_mode="highpass", bp_order=4,
    operating_band=(400.0, 3500.0), q=0.25, win_sec=0.5, adaptive_q_enable=False, adaptive_q_min=0.10,
    adaptive_q_alpha=0.95, median_frames=0, eps=1e-9, noise_psd_max_ratio=1.0, pre_smooth_frames=0, ema_up=0.6,
    ema_down=0.95, detector_use_noise_norm=True, detector_noise_norm_mode="log_sub",
    suppressor_bypass=False, classifier_only_mode=False, process_dtype="float32")


def resolve(params):
    """Reference config precedence: flat > params['suppressor'] > dataclass defaults;
    detector lookups: params['detector'] > cfg attribute > default."""
    params = dict(params)
    fs = int(params.get("sample_rate", params.get("fs", 11162)))
    sup = params.get("suppressor")
    if isinstance(sup, dict):
        params = {**sup, **params}
    det = params.get("detector") if isinstance(params.get("detector"), dict) else {}
    if "operating_band" not in params and params.get("fmin") is not None and params.get("fmax") is not None:
        params["operating_band"] = (float(params["fmin"]), float(params["fmax"]))
    cfg = dict(_CFG_DEFAULTS)
    cfg["fs"] = fs
    for k, v in params.items():
        if k in cfg and k != "fs":
            cfg[k] = v
    cfg["operating_band"] = (float(cfg["operating_band"][0]), float(cfg["operating_band"][1]))

    def dget(name, default=None):
        if name in det:
            return det[name]
        if name in cfg:
            return cfg[name]
        return default

    for flag, bad in (("process_dtype", "float64"),):
        if cfg[flag] == bad:
            raise NotImplementedError(f"oracle: {flag}={bad!r} not restated")
    for name in ("peak_features_enable", "flux_modes_winsor_enable", "td_envelope_features_enable"):
        if bool(dget(name, False)):
            raise NotImplementedError(f"oracle: detector.{name} not restated")
    if str(dget("td_input_mode", "default")).lower() != "default":
        raise NotImplementedError("oracle: td_input_mode != 'default' not restated")
    return cfg, det, dget


def _bin_range(mask):
    idx = np.flatnonzero(mask)
    if idx.size == 0:
        return 1, 0
    assert idx[-1] - idx[0] + 1 == idx.size
    return int(idx[0]), int(idx[-1])


def make_params(params):
    cfg, det, dget = resolve(params)
    f32 = np.float32
    P = OrcParams()
    fs, n_fft, hop = cfg["fs"], int(cfg["n_fft"]), int(cfg["hop"])
    P.fs, P.n_fft, P.hop = fs, n_fft, hop
    # rain_signal_processor.py:827 -- freqs in the work dtype (float32)
    freqs = np.asarray(np.fft.rfftfreq(n=n_fft, d=1.0 / fs), dtype=np.float32)
    op_lo, op_hi = cfg["operating_band"]
    band = (freqs >= op_lo) & (freqs <= op_hi)
    P.band_lo, P.band_hi = _bin_range(band)
    mode_bands = dget("mode_bands", None)
    if mode_bands is None:
        raise AttributeError("Missing required detector param: mode_bands")
    mode_bands = tuple((float(a), float(b)) for a, b in mode_bands)
    if len(mode_bands) < 4:
        raise ValueError("Fixed-band rain decision requires at least 4 mode bands")
    P.n_modes = len(mode_bands)
    fb = freqs[band]
    f64 = freqs.astype(np.float64)
    weights = dget("mode_weights", None)
    for i, (lo, hi) in enumerate(mode_bands):
        P.mode_lo[i], P.mode_hi[i] = _bin_range((f64 >= lo) & (f64 <= hi))
        P.mode_in_band_lo[i], P.mode_in_band_hi[i] = _bin_range((fb >= lo) & (fb <= hi))
        P.mode_weight[i] = float(weights[i]) if weights is not None else 1.0
    # tracker constants (rain_signal_processor.py:562-567, 683-684)
    frames_per_sec = float(fs) / float(hop)
    W = max(10, int(cfg["win_sec"] * frames_per_sec))
    eta = float(np.clip(float(2.0 / max(W + 1, 2)), 1e-4, 1.0))
    scale_alpha = float(cfg["ema_down"])
    step_floor = float(max(cfg["eps"], 1e-9))
    maxr = float(cfg["noise_psd_max_ratio"])
    maxr = 1.0 if not np.isfinite(maxr) else float(np.clip(maxr, 0.0, 1.0))
    q = float(cfg["q"])
    P.trk_eta, P.trk_scale_alpha, P.trk_one_minus_alpha = f32(eta), f32(scale_alpha), f32(1.0 - scale_alpha)
    P.trk_step_floor, P.trk_q, P.trk_neg_one_minus_q, P.trk_maxr = f32(step_floor), f32(q), f32(-(1.0 - q)), f32(maxr)
    P.ema_up, P.ema_down = float(cfg["ema_up"]), float(cfg["ema_down"])
    P.adaptive_q = int(bool(cfg["adaptive_q_enable"]))            # rain_signal_processor.py:570-576
    P.aq_base = q
    P.aq_min = float(np.clip(float(cfg["adaptive_q_min"]), 1e-4, q))
    P.aq_alpha = float(np.clip(float(cfg["adaptive_q_alpha"]), 0.0, 1.0))
    P.bypass_classifier = int(bool(dget("bypass_classifier", False)))   # rain_signal_processor.py:846-857
    P.pre_smooth_frames = int(cfg["pre_smooth_frames"] or 0)      # rain_signal_processor.py:690-692
    P.median_frames = int(cfg["median_frames"] or 0)              # :717-719
    P.warmup_need = max(10, W // 2)
    P.eps_f32 = f32(cfg["eps"])
    P.detector_use_noise_norm = int(bool(dget("detector_use_noise_norm", True)))
    P.norm_ratio_db = int(str(cfg["detector_noise_norm_mode"]).lower() == "ratio_db")
    # flux baseline (rain_frame_classifier.py:52-58, 438-443, 845-859)
    eps = float(dget("eps", 1e-9))
    qp = float(np.clip(float(dget("mode_flux_norm_q", 20.0)), 0.0, 100.0))
    norm_min = max(float(dget("mode_flux_norm_min", 1.0)), eps)
    fps = float(dget("sample_rate", dget("fs", 11162))) / max(float(dget("hop", 128)), 1.0)
    sps = float(max(fps, 1e-6))
    Wb = max(3, int(round(float(dget("mode_flux_norm_win_sec", 0.5)) * sps)))
    bl_eta = float(np.clip(2.0 / max(Wb + 1, 2), 1e-4, 1.0))
    P.bl_q = float(np.clip(qp, 0.0, 100.0)) / 100.0
    P.bl_eta = bl_eta
    P.bl_scale_alpha = float(np.clip(1.0 - bl_eta, 0.0, 0.9999))
    P.bl_floor = float(max(norm_min, 1e-12))
    P.norm_enable = int(bool(dget("mode_flux_norm_enable", True)))
    P.norm_min_f32 = f32(norm_min)
    legacy12 = float(dget("new_rain_mode12_flux_min", 2.6))
    P.thr_primary = f32(float(dget("new_rain_primary_flux_min", 1.8)))
    P.thr_m1 = f32(float(dget("new_rain_mode1_flux_min", legacy12)))
    P.thr_m2 = f32(float(dget("new_rain_mode2_flux_min", legacy12)))
    P.thr_m3 = f32(float(dget("new_rain_mode3_flux_min", 3.0)))
    P.min_support = int(dget("new_rain_min_support_count", 2))
    P.td_gate_thr = f32(float(dget("td_gate_threshold", 2.5)))
    ku = dget("td_kurtosis_upper_threshold", None)
    P.has_kurt_upper = int(ku is not None)
    P.kurt_upper = f32(float(ku)) if ku is not None else f32(0)
    P.noise_hi = f32(float(dget("noise_hi", 0.80)))
    P.mode_flux_noise_max = f32(max(float(dget("mode_flux_noise_max", 1.5)), 0.0))
    # TD prefilter (rain_signal_processor.py:347-364, rain_frame_classifier.py:397-403, 472-481)
    td_mode = str(dget("td_prefilter_mode", dget("pre_filter_mode", "none"))).lower()
    sos = None
    if bool(dget("td_apply_input_prefilter", True)) and td_mode not in ("", "none"):
        nyq = 0.5 * fs
        if td_mode == "bandpass":
            lo = np.clip(float(op_lo), 1e-3, nyq * 0.999)
            hi = np.clip(float(op_hi), lo + 1e-3, nyq * 0.999)
            sos = spsig.butter(int(cfg.get("bp_order", cfg["hp_order"])), [lo / nyq, hi / nyq],
                               btype="bandpass", output="sos")
        elif td_mode == "highpass" and cfg["hp_cutoff_hz"] > 0:
            sos = spsig.butter(cfg["hp_order"], np.clip(cfg["hp_cutoff_hz"] / nyq, 1e-4, 0.9999),
                               btype="highpass", output="sos")
    if sos is None:
        P.n_sos, P.padlen = 0, 0
    else:
        ns = sos.shape[0]
        assert ns <= MAXS
        ntaps = 2 * ns + 1
        ntaps -= min((sos[:, 2] == 0).sum(), (sos[:, 5] == 0).sum())
        P.n_sos, P.padlen = ns, int(3 * ntaps)
        zi = spsig.sosfilt_zi(sos)
        for s in range(ns):
            for j in range(6):
                P.sos[s][j] = float(sos[s, j])
            P.zi[s][0], P.zi[s][1] = float(zi[s, 0]), float(zi[s, 1])
    P.eps_f64 = eps
    P.blk_len = int(max(1, int(dget("td_block_energy_len", 8))))
    bh = dget("td_block_energy_hop", None)
    P.blk_hop = max(1, int(bh)) if bh is not None else P.blk_len
    P.blk_post_pre = int(dget("td_block_energy_post_pre_blocks", 4))
    P.blk_smooth = int(bool(dget("td_block_energy_smooth_enable", True)))
    low = dget("raw_spectral_low_band", (50.0, 200.0))
    rain = dget("raw_spectral_rain_band", (400.0, 800.0))
    P.low_lo, P.low_hi = _bin_range((f64 >= max(float(low[0]), eps)) & (f64 < float(low[1])))
    P.rain_lo, P.rain_hi = _bin_range((f64 >= float(rain[0])) & (f64 <= float(rain[1])))
    P.rolloff_fraction = float(dget("raw_spectral_rolloff_fraction", 0.85))
    P.suppressor_bypass = int(bool(cfg["suppressor_bypass"]))
    window = spsig.get_window("hann", n_fft, fftbins=True).astype(np.float64)
    return P, window, freqs, cfg, band


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def process(x_f32, params, want="all"):
    """Run the oracle on one float32 clip.  Returns a dict of arrays (time-major planes)."""
    L = lib()
    P, window, freqs, cfg, band = make_params(params)
    x = np.ascontiguousarray(np.asarray(x_f32, dtype=np.float32).reshape(-1))
    n = x.size
    T = int(L.orc_num_frames(n, P.hop))
    F = P.n_fft // 2 + 1
    K = P.band_hi - P.band_lo + 1
    M = P.n_modes
    full = want == "all"
    o = {
        "frame_class": np.zeros(T, np.int8), "rain_conf": np.zeros(T, np.float32),
        "noise_conf": np.zeros(T, np.float32), "N2_band": np.zeros((T, K), np.float32),
        "td": np.zeros((5, T), np.float32), "gate": np.zeros(T, np.uint8),
        "norm_flux": np.zeros((M, T), np.float32), "score": np.zeros(T, np.float32),
    }
    if full:
        o.update({
            "S": np.zeros((T, F, 2), np.float32), "P_band": np.zeros((T, K), np.float32),
            "N1_band": np.zeros((T, K), np.float32), "Nlag_band": np.zeros((T, K), np.float32),
            "D_band": np.zeros((T, K), np.float32), "mode_flux": np.zeros((M, T), np.float32),
            "flux_modes": np.zeros(T, np.float32), "baseline": np.zeros((M + 1, T), np.float32),
            "x_td": np.zeros(n, np.float32), "raw": np.zeros((21, T), np.float32),
        })
    out = OrcOut()
    for k, v in o.items():
        setattr(out, k, v.ctypes.data)
    rc = L.orc_process(C.byref(P), _ptr(window), _ptr(freqs), _ptr(x), n, C.byref(out))
    if rc != 0:
        raise RuntimeError(f"oracle failed rc={rc}")
    if full:
        o["S"] = o["S"].view(np.complex64).reshape(T, F)
    o["times"] = np.asarray((np.arange(T) * P.hop).astype(int) / float(P.fs), dtype=np.float32)
    o["freqs"] = freqs
    o["band_mask"] = band
    o["_cfg"] = cfg
    o["_P"] = P
    return o


def noise_floor_stats(N2_band, eps):
    """mean/median of 10*log10(noise_psd[band] + eps) (rain_signal_processor.py:1292-1298).

    The array is laid out (K, T) C-contiguous like ``noise_psd[band_mask]`` in the reference,
    and numpy does the float32 reductions, so the summation order is the reference's."""
    L = lib()
    NB = np.ascontiguousarray(N2_band.T)          # (K, T)
    db = np.empty_like(NB)
    L.orc_noise_db(_ptr(NB), NB.size, np.float32(eps), _ptr(db))
    return float(np.mean(db)), float(np.median(db))


def run(audio, params):
    """Oracle twin of RainDetectorProcessor.run (metrics, state) without timing keys."""
    params = dict(params)
    o = process(np.asarray(audio, dtype=np.float32), params,
                want="all" if params.get("keep_state_debug") or params.get("keep_state_spectra") else "core")
    fc = o["frame_class"]
    is_rain = fc == 2
    min_frames = max(1, int(params.get("clip_rain_min_frames", 1)))
    count = int(np.sum(is_rain))
    frac = float(np.mean(is_rain)) if is_rain.size else 0.0
    med = float(np.median(o["rain_conf"][is_rain])) if count > 0 else 0.0
    abundance = float(np.clip(count / float(max(2 * min_frames, 1)), 0.0, 1.0))
    metrics = {
        "rain_frame_fraction": frac, "clip_rain_fraction": frac, "rain_frame_count": count,
        "clip_is_rain": bool(count >= min_frames), "clip_rain_conf": float(max(med, abundance)),
        "median_rain_conf": med, "clip_rain_min_frames": min_frames,
    }
    mean_db, med_db = noise_floor_stats(o["N2_band"], o["_cfg"]["eps"])
    metrics["mean_noise_floor_db"] = mean_db
    metrics["median_noise_floor_db"] = med_db
    state = dict(o)
    state["event_idx"] = np.flatnonzero(is_rain).astype(np.int32)
    return metrics, state


def process_batch_i16(pcm_list, params, n_threads=1):
    """Threaded batch driver (CPU baseline timing): returns (frame_class list, rain counts)."""
    L = lib()
    P, window, freqs, cfg, band = make_params(params)
    lens = np.array([len(p) for p in pcm_list], dtype=np.int64)
    offs = np.zeros(len(pcm_list) + 1, np.int64)
    offs[1:] = np.cumsum(lens)
    pcm = np.ascontiguousarray(np.concatenate([np.asarray(p, np.int16) for p in pcm_list]))
    Ts = 1 + lens // P.hop
    foffs = np.zeros(len(pcm_list) + 1, np.int64)
    foffs[1:] = np.cumsum(Ts)
    fc = np.zeros(int(foffs[-1]), np.int8)
    cnt = np.zeros(len(pcm_list), np.int32)
    rc = L.orc_process_batch_i16(C.byref(P), _ptr(window), _ptr(freqs), _ptr(pcm), _ptr(offs),
                                 len(pcm_list), _ptr(fc), _ptr(foffs), _ptr(cnt), int(n_threads))
    if rc != 0:
        raise RuntimeError(f"oracle batch failed rc={rc}")
    return [fc[foffs[i]:foffs[i + 1]] for i in range(len(pcm_list))], cnt
