"""GPU-backed twins of the reference's spectral noise engine and its framework-facing processor.

  * ``SpectralNoiseProcessor``  -- edge/rain_signal_processor.py:257-1198 (``setup`` / ``process``)
  * ``RainDetectorProcessor``   -- edge/rain_signal_processor.py:1205-1344 (``run``), plus the additive
                                   ``run_batch`` hook the batched orchestrator uses
  * ``NoiseProcessorConfig`` / ``build_noise_config`` re-exported from ..config

Every array is computed by the CUDA kernels behind include/apt_b200.h; this module only validates
inputs, resolves parameters and packages results into the reference's dictionaries.
"""
from __future__ import annotations

import json
import time
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from ..config import DetectorView, NoiseProcessorConfig, build_noise_config, validate_config
from ..engine import BatchEngine
from ..processors import BaseProcessor
from .rain_frame_classifier import FrameClass

RAW_SPECTRAL_FEATURE_NAMES = (
    "raw_spectral_centroid_hz", "raw_spectral_bandwidth_hz", "raw_low_freq_ratio",
    "raw_rain_band_ratio", "raw_mode_band_ratio_0", "raw_mode_band_ratio_1",
    "raw_mode_band_ratio_2", "raw_mode_band_ratio_3", "raw_mode_band_ratio_4",
    "raw_mode_band_entropy", "raw_mode_band_std", "raw_mode_band_max_ratio",
    "raw_spectral_flatness", "raw_spectral_rolloff_hz", "raw_dominant_freq_hz",
    "raw_frame_energy", "raw_cepstrum_coeff_0", "raw_cepstrum_coeff_1",
    "raw_cepstrum_coeff_2", "raw_cepstrum_coeff_3", "raw_cepstrum_coeff_4")
TD_FEATURE_ROWS = ("td_crest_factor", "td_kurtosis", "td_block_energy_crest",
                   "td_block_peak_width_50", "td_block_post_pre_energy_ratio")

__all__ = ["NoiseProcessorConfig", "build_noise_config", "SpectralNoiseProcessor",
           "RainDetectorProcessor", "FrameClass"]


class SpectralNoiseProcessor:
    """STFT -> rain-frame detection -> noise-PSD estimation, on the GPU, for one or many clips."""

    def __init__(self, config: Optional[NoiseProcessorConfig] = None, *, device: int = 0, fft_f64: bool = True):
        self.cfg = config
        self._device = device
        self._fft_f64 = fft_f64
        self._engine: Optional[BatchEngine] = None
        self._engine_key = None
        self._is_setup = config is not None
        if self._is_setup:
            validate_config(self.cfg)

    def setup(self, params: Dict[str, Any]):
        if self._is_setup:
            return
        sr = int(params.get("sample_rate", params.get("fs", 11162)))
        self.cfg = build_noise_config(sample_rate=sr, params=params)
        validate_config(self.cfg)
        self._is_setup = True

    # -- engine management -------------------------------------------------------------------
    def _get_engine(self, sr: int, clip_rain_min_frames: int = 1) -> BatchEngine:
        key = (int(sr), int(clip_rain_min_frames))
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            self._engine = BatchEngine(self.cfg, sr, device=self._device,
                                       clip_rain_min_frames=clip_rain_min_frames, fft_f64=self._fft_f64)
            self._engine_key = key
        return self._engine

    def __getstate__(self):  # engines own GPU resources: never pickled (ProcessPool boundary)
        d = dict(self.__dict__)
        d["_engine"] = None
        d["_engine_key"] = None
        return d

    # -- main API ---------------------------------------------------------------------------
    def process(self, x: np.ndarray, sr: Optional[int] = None) -> Dict[str, Any]:
        return self.process_batch([x], sr=sr)[0]

    def process_batch(self, clips: Sequence[np.ndarray], sr: Optional[int] = None,
                      clip_rain_min_frames: int = 1, with_stats: bool = False,
                      with_events: bool = True) -> List[Dict[str, Any]]:
        if self.cfg is None:
            self.setup({"sample_rate": sr or 11162})
        cfg = self.cfg
        if sr is None:
            sr = cfg.fs
        want_audio = bool(cfg.compute_output_audio)
        dv = DetectorView(cfg)
        keep_debug = bool(cfg.return_debug) or bool(cfg.debug_enable)
        keep_det = bool(cfg.return_detector_debug) or bool(cfg.debug_enable)
        # the features payload carries the detector-side dump when feature_dump_level > 0: its arrays are needed then
        need_det = keep_det or (bool(cfg.dump_features) and int(dv.get("feature_dump_level", 0)) > 0)
        bypass_cls = bool(dv.get("bypass_classifier", False))
        keep_spectra = bool(cfg.return_spectra)
        keep_noise = bool(cfg.return_noise_psd)
        keep_filt = bool(cfg.return_filtered_audio)
        want = []
        suppress = not (bool(cfg.suppressor_bypass) or bool(cfg.classifier_only_mode))
        if keep_debug:
            want += ["det_noise_psd", "det_noise_lag", "noise_psd"]
            if suppress:
                want += ["G", "ratio_med"]
                if bool(cfg.snr_gating_enable):
                    want += ["snr_mode", "snr_gate"]
        if keep_noise and "noise_psd" not in want:
            want.append("noise_psd")
        peaks = need_det and bool(dv.get("peak_features_enable", False))
        if peaks:
            want += ["peak_ratio", "peak_gate_score", "peak_valid_count", "peak_count_by_mode"]
        if need_det and not bypass_cls:
            want += ["norm_flux", "score", "td", "gate"]
            if dv.get("raw_spectral_shape_enable", True):
                want.append("raw")
        if keep_spectra:
            want.append("S")
            if suppress:
                want += ["S_hat"] + ([] if "G" in want else ["G"])
        if keep_filt or want_audio:
            want.append("x_td")
        if want_audio and suppress:
            want += [w for w in ("S", "G", "S_hat", "y") if w not in want]
        arrays = []
        for x in clips:
            a = np.asarray(x)
            arrays.append(a.reshape(-1) if a.dtype == np.int16 else np.asarray(a, dtype=np.float32).reshape(-1))
        eng = self._get_engine(int(sr), clip_rain_min_frames)
        if not want:
            # default flags (labels, confidences, clip statistics only): the pipelined host path -- pinned staging
            # ring, host->device copies overlapped with compute, results straight into the caller's arrays
            plan, out = eng.run_host_clips(arrays, event_idx=with_events)
            self.last_host_call_s = eng.last_host_call_s
            return self._package_core(plan, out, cfg, sr, eng.rp, with_stats)
        plan, out = eng.run_clips(arrays, want)
        rp = eng.rp
        band = rp.band_mask
        results = []
        for c in range(plan.n_clips):
            f0, f1 = int(plan.frame_off[c]), int(plan.frame_off[c + 1])
            s0, s1 = int(plan.sample_off[c]), int(plan.sample_off[c + 1])
            T = f1 - f0
            times = np.asarray((np.arange(T) * int(cfg.hop)).astype(int) / float(sr), dtype=np.float32)
            res: Dict[str, Any] = {
                "frame_class": out["frame_class"][f0:f1].copy(),
                "freqs": rp.freqs.copy(),
                "times": times,
                "rain_conf": out["rain_conf"][f0:f1].copy(),
                "noise_conf": out["noise_conf"][f0:f1].copy(),
            }
            if need_det and bypass_cls:
                # bypass_classifier (rain_signal_processor.py:846-857): the reference's placeholder dictionary
                fcn = res["frame_class"]
                dd = {"frame_class": fcn, "frame_class_name": np.array(["noise"] * T, dtype=object),
                      "rain_score_raw": np.zeros(T, dtype=np.float32), "is_rain_raw": np.zeros(T, dtype=bool),
                      "onset_mask": np.zeros(T, dtype=bool), "onset_indices": np.zeros(0, dtype=np.int32)}
            else:
                dd = self._det_debug(out, f0, f1, rp, dv) if need_det else None
            if keep_det:
                res["det_debug"] = dd
            if bool(cfg.dump_features):
                res["features"] = self._features(cfg, times, res["frame_class"], res["rain_conf"], res["noise_conf"], dd)
            if keep_debug:
                res["debug"] = self._debug(out, f0, f1, rp, dv, times)
            if keep_filt:
                xf = out["x_td"][s0:s1].copy()
                res["x_filt"] = xf
                if bool(cfg.classifier_only_mode):
                    res["y"] = xf
                else:
                    # suppressed waveform (rain_signal_processor.py:1113-1128); bypass returns the prefiltered input
                    yh = (out["y"][s0:s1].copy() if suppress else xf.copy()) if want_audio else None
                    res["y"] = yh
                    res["y_suppressed"] = yh
            if keep_spectra:
                S = np.asfortranarray(out["S"][f0:f1].view(np.complex64).reshape(T, rp.F).T)
                res["S"] = S
                if suppress:
                    res["S_hat"] = np.asfortranarray(out["S_hat"][f0:f1].view(np.complex64).reshape(T, rp.F).T)
                else:
                    res["S_hat"] = S.copy()
            if keep_noise and not bool(cfg.classifier_only_mode):
                res["noise_psd"] = self._embed(out["noise_psd"][f0:f1], rp)
            if with_stats:
                res["_clip_stats"] = out["clip_stats"][c].copy()
                n = int(out["event_count"][c])
                res["_event_idx"] = out["event_idx"][f0:f0 + n].copy()
            results.append(res)
        return results

    _times_cache: Dict[tuple, np.ndarray] = {}

    @classmethod
    def _times(cls, T: int, hop: int, sr: float) -> np.ndarray:
        """librosa.frames_to_time as float32 (rain_signal_processor.py:828).  One READ-ONLY array per (T, hop, sr): every
        clip of that length gets the same object (a private copy per clip costs 0.2 MB and 20 us for a ten-minute clip,
        20 ms per 1 000-clip batch); callers that want to edit it copy it."""
        key = (int(T), int(hop), float(sr))
        t = cls._times_cache.get(key)
        if t is None:
            if len(cls._times_cache) > 64:
                cls._times_cache.clear()
            t = np.asarray((np.arange(T) * int(hop)).astype(int) / float(sr), dtype=np.float32)
            t.setflags(write=False)
            cls._times_cache[key] = t
        return t

    @staticmethod
    def _features(cfg, times, fc, rc, nc, det_debug=None) -> Dict[str, Any]:
        """`dump_features` payload (rain_signal_processor.py:723-787): the five per-frame arrays (plus the detector-side
        feature dump, which is empty at the supported feature_dump_level = 0), decimated along the frame axis by `feature_decim`."""
        step = max(1, int(getattr(cfg, "feature_decim", 1)))

        def dec(v):
            if step <= 1 or v is None:
                return v
            if isinstance(v, np.ndarray):
                return v if v.ndim == 0 else v[..., ::step]
            if isinstance(v, (list, tuple)):
                return v[::step]
            return v

        f = {"frame_times": dec(np.asarray(times, dtype=np.float32)), "frame_class": dec(np.asarray(fc)),
             "is_rain": dec(np.asarray(fc) == FrameClass.RAIN), "rain_conf": dec(np.asarray(rc, dtype=np.float32)),
             "noise_conf": dec(np.asarray(nc, dtype=np.float32))}
        if isinstance(det_debug, dict):
            # the detector-side dump wins when it is a dict (:771-776) -- and under feature_dump_level = 0 it is an EMPTY
            # dict, so the reference's payload is the five arrays above; only without it are the detector's arrays passed on
            fd = det_debug.get("feature_dump", None)
            if isinstance(fd, dict):
                for k, v in fd.items():
                    f[k] = dec(v)
                return f
            for k, v in det_debug.items():
                if k != "feature_dump":
                    f[k] = dec(v)
        return f

    def _package_core(self, plan, out, cfg, sr, rp, with_stats) -> List[Dict[str, Any]]:
        """Result dictionaries of the default-flags path: per-clip views of the batch's (caller-owned) host arrays."""
        results = []
        fo = plan.frame_off
        fc, rc, nc = out["frame_class"], out["rain_conf"], out["noise_conf"]
        ev = out.get("event_idx")
        for c in range(plan.n_clips):
            f0, f1 = int(fo[c]), int(fo[c + 1])
            res: Dict[str, Any] = {"frame_class": fc[f0:f1], "freqs": rp.freqs.copy(),
                                   "times": self._times(f1 - f0, cfg.hop, sr),
                                   "rain_conf": rc[f0:f1], "noise_conf": nc[f0:f1]}
            if bool(cfg.dump_features):
                res["features"] = self._features(cfg, res["times"], res["frame_class"], res["rain_conf"], res["noise_conf"])
            if with_stats:
                res["_clip_stats"] = out["clip_stats"][c]
                res["_event_idx"] = ev[f0:f0 + int(out["event_count"][c])] if ev is not None else None
            results.append(res)
        return results

    @staticmethod
    def _embed(plane_tk: np.ndarray, rp, fill: float = 0.0) -> np.ndarray:
        """[T][K] band plane -> (F, T) float32 with `fill` off-band (Fortran order like the reference)."""
        T = plane_tk.shape[0]
        full = np.full((rp.F, T), fill, dtype=np.float32, order="F")
        full[rp.c.band_lo:rp.c.band_hi + 1, :] = plane_tk.T
        return full

    def _det_debug(self, out, f0, f1, rp, dv) -> Dict[str, Any]:
        nf = out["norm_flux"][:, f0:f1]
        gate = out["gate"][f0:f1].astype(bool)
        gs = gate.astype(np.float32)
        score = out["score"][f0:f1].copy()
        d: Dict[str, Any] = {
            "mode_flux_score": score,
            "mode_flux_score_gated": score * gs,
            "primary_mode_flux": nf[0].copy(),
            "support_mode_flux_1": nf[1].copy(),
            "support_mode_flux_2": nf[2].copy(),
            "support_mode_flux_3": nf[3].copy(),
            "support_mode_flux_4": nf[4].copy() if nf.shape[0] > 4 else np.zeros_like(nf[0]),
            "rain_conf": out["rain_conf"][f0:f1].copy(),
            "noise_conf": out["noise_conf"][f0:f1].copy(),
            "frame_class": out["frame_class"][f0:f1].copy(),
            "td_gate_mask": gate,
            "td_gate_threshold": float(dv.get("td_gate_threshold", 2.5)),
            "td_kurtosis_upper_threshold": dv.get("td_kurtosis_upper_threshold", None),
            "peak_features_enable": "peak_ratio" in out,
        }
        if "peak_ratio" in out:
            d["peak_ratio"] = out["peak_ratio"][f0:f1].copy()
            d["peak_gate_score"] = out["peak_gate_score"][f0:f1].copy()
            d["peak_valid_count"] = out["peak_valid_count"][f0:f1].copy()
            d["peak_count_by_mode"] = out["peak_count_by_mode"][:, f0:f1].copy()
        for i in range(1, 4):
            d[f"support_mode_flux_{i}_gated"] = d[f"support_mode_flux_{i}"] * gs
        d["primary_mode_flux_gated"] = d["primary_mode_flux"] * gs
        for i, name in enumerate(TD_FEATURE_ROWS):
            d[name] = out["td"][i, f0:f1].copy()
        raw_on = "raw" in out
        for i, name in enumerate(RAW_SPECTRAL_FEATURE_NAMES):
            d[name] = out["raw"][i, f0:f1].copy() if raw_on else np.zeros(f1 - f0, dtype=np.float32)
        # soft TD label (assign_td_soft_label, rain_frame_classifier.py:85-110, :618-629): votes of crest factor and kurtosis
        T = f1 - f0
        votes = np.zeros(T, dtype=np.int32)
        if bool(dv.get("td_soft_enable", False)):
            votes += (d["td_crest_factor"] >= float(dv.get("td_soft_crest_factor_min", 4.0))).astype(np.int32)
            votes += (d["td_kurtosis"] >= float(dv.get("td_soft_kurtosis_min", 6.0))).astype(np.int32)
            d["td_soft_score"] = votes.astype(np.float32) / 2.0
            d["td_soft_label"] = votes >= int(dv.get("td_soft_min_positive_votes", 2))
        else:
            d["td_soft_score"] = np.zeros(T, dtype=np.float32)
            d["td_soft_label"] = np.zeros(T, dtype=bool)
        d["td_vote_count"] = votes
        # sparse-dump bookkeeping (:945-967) and the scalar echoes of the detector's configuration (:1022-1045)
        gate_feature = str(dv.get("feature_dump_sparse_gate_feature", "td_block_energy_crest"))
        sparse_on = bool(dv.get("feature_dump_sparse_enable", False))
        sparse_thr = float(dv.get("feature_dump_sparse_gate_threshold", 3.5))
        if sparse_on:
            src = d["td_crest_factor"] if gate_feature == "td_crest_factor" else d["td_block_energy_crest"]
            mask = np.nan_to_num(src, nan=0.0, posinf=0.0, neginf=0.0) > sparse_thr
        else:
            mask = np.ones(T, dtype=bool)
        d["raw_spectral_dump_mask"] = mask
        d["raw_spectral_dump_mask_fraction"] = float(np.mean(mask.astype(np.float32))) if T > 0 else 0.0
        d["sparse_frame_idx"] = np.flatnonzero(mask).astype(np.int32)
        blk_len = int(dv.get("td_block_energy_len", 8))
        blk_hop = dv.get("td_block_energy_hop", None)
        d.update({
            "td_block_energy_len": blk_len,
            "td_block_energy_hop": int(blk_hop) if blk_hop is not None else None,
            "td_block_energy_post_pre_blocks": int(dv.get("td_block_energy_post_pre_blocks", 4)),
            "td_block_energy_smooth_enable": bool(dv.get("td_block_energy_smooth_enable", True)),
            "feature_dump_dense_enable": bool(dv.get("feature_dump_dense_enable", True)),
            "feature_dump_sparse_enable": sparse_on,
            "feature_dump_clip_summary_enable": bool(dv.get("feature_dump_clip_summary_enable", False)),
            "feature_dump_sparse_gate_feature": gate_feature,
            "feature_dump_sparse_gate_threshold": sparse_thr,
            "raw_spectral_shape_enable": bool(dv.get("raw_spectral_shape_enable", True)),
            "raw_spectral_uses_raw_power": True,
            "td_apply_input_prefilter": bool(dv.get("td_apply_input_prefilter", True)),
            "td_prefilter_mode": str(dv.get("td_prefilter_mode", dv.get("pre_filter_mode", "none"))).lower(),
            "clip_spectral_occupancy_enable": bool(dv.get("clip_spectral_occupancy_enable", False)),
        })
        # detector-side feature dump (rain_frame_classifier.py:1096-1162): empty at feature_dump_level = 0, else the dense
        # per-frame arrays and / or the raw spectral features at the sparse frames (the clip summary needs
        # clip_spectral_occupancy_enable, which is refused)
        fd: Dict[str, Any] = {}
        if int(dv.get("feature_dump_level", 0)) > 0:
            if d["feature_dump_dense_enable"]:
                for k in ("primary_mode_flux", "support_mode_flux_1", "support_mode_flux_2", "support_mode_flux_3",
                          "support_mode_flux_4", "td_block_energy_crest", "td_block_peak_width_50",
                          "td_block_post_pre_energy_ratio", "td_gate_mask"):
                    fd[k] = d[k]
                if bool(dv.get("feature_dump_include_frame_class", True)):
                    fd["frame_class"] = d["frame_class"]
                if bool(dv.get("feature_dump_include_td_soft", False)):
                    for k in ("td_crest_factor", "td_kurtosis", "td_vote_count", "td_soft_score"):
                        fd[k] = d[k]
            if sparse_on:
                idx = d["sparse_frame_idx"]
                fd["sparse_frame_idx"] = idx
                basic = ("raw_spectral_centroid_hz", "raw_rain_band_ratio", "raw_spectral_rolloff_hz")
                inc_basic = bool(dv.get("feature_dump_include_raw_spectral_basic", False))
                if bool(dv.get("feature_dump_include_raw_spectral_frame_features", True)):
                    for name in RAW_SPECTRAL_FEATURE_NAMES:
                        if name in basic and not inc_basic:
                            continue
                        fd["sparse_" + name] = d[name][idx]
                elif inc_basic:
                    for name in basic:
                        fd["sparse_" + name] = d[name][idx]
        d["feature_dump"] = fd
        return d

    def _debug(self, out, f0, f1, rp, dv, times) -> Dict[str, Any]:
        cfg = self.cfg
        fc = out["frame_class"][f0:f1]
        use = fc == FrameClass.NOISE
        return {
            "detector_params": dict(cfg.detector or {}),
            "suppressor_params": dict(cfg.suppressor or {}),
            "times_s": times,
            "freqs": rp.freqs.copy(),
            # (no detector pass under bypass_classifier: the reference leaves both None)
            "detector_noise_psd": None if bool(dv.get("bypass_classifier", False)) else self._embed(out["det_noise_psd"][f0:f1], rp),
            "detector_noise_psd_lag": None if bool(dv.get("bypass_classifier", False)) else self._embed(out["det_noise_lag"][f0:f1], rp),
            "detector_use_noise_norm": bool(dv.get("detector_use_noise_norm", True)),
            "detector_noise_norm_mode": str(cfg.detector_noise_norm_mode).lower(),
            "suppressor_bypass": bool(cfg.suppressor_bypass),
            "classifier_only_mode": bool(cfg.classifier_only_mode),
            "use_for_noise_psd": use,
            "is_rain_for_psd": ~use,
            "noise_psd": self._embed(out["noise_psd"][f0:f1], rp),
            "G": (self._embed(out["G"][f0:f1], rp, fill=1.0) if "G" in out
                  else np.ones((rp.F, f1 - f0), dtype=np.float32, order="F")),
            "np_ratio_median_t": (out["ratio_med"][f0:f1].copy() if "ratio_med" in out
                                  else np.zeros(f1 - f0, dtype=np.float32)),
            "use_lagged_noise_psd": bool(cfg.use_lagged_noise_psd),
            "operating_band": (float(cfg.operating_band[0]), float(cfg.operating_band[1])),
            "band_mask": rp.band_mask.copy(),
            "pre_filter_mode": str(cfg.pre_filter_mode).lower(),
            "pre_filter_band": (float(cfg.operating_band[0]), float(cfg.operating_band[1])),
            "noise_psd_max_ratio": float(cfg.noise_psd_max_ratio),
            "td_soft_enable": bool(dv.get("td_soft_enable", False)),
            "bypass_classifier": bool(dv.get("bypass_classifier", False)),
            "features_available": bool(cfg.dump_features),
            # frame SNR and gate of the spectral SNR gating (:1050-1077); None when it is off, as in the reference
            "snr_mode": out["snr_mode"][f0:f1].copy() if "snr_mode" in out else None,
            "snr_gate": out["snr_gate"][f0:f1].copy() if "snr_gate" in out else None,
        }


class RainDetectorProcessor(BaseProcessor):
    """Framework-facing processor for rain-frame detection (GPU).  ``run`` keeps the reference's
    per-file contract; ``run_batch`` processes a whole batch of clips in one GPU pass."""

    def __init__(self, name: str = "rain_detector", *, device: int = 0, fft_f64: bool = True):
        self.name = name
        self._device = device
        self._fft_f64 = fft_f64
        self._proc_cache: Dict[str, SpectralNoiseProcessor] = {}

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_proc_cache"] = {}
        return d

    def _params_cache_key(self, params: Dict[str, Any]) -> str:
        try:
            return json.dumps(params, sort_keys=True, default=str)
        except Exception:
            return repr(sorted(params.items(), key=lambda kv: kv[0]))

    def _prepare(self, params: Dict[str, Any]):
        p = dict(params)
        keep_audio = bool(p.get("keep_state_audio", False))
        keep_spectra = bool(p.get("keep_state_spectra", False))
        keep_debug = bool(p.get("keep_state_debug", False))
        p.setdefault("compute_output_audio", keep_audio)
        p.setdefault("return_filtered_audio", keep_audio)
        p.setdefault("return_spectra", keep_spectra)
        p.setdefault("return_debug", keep_debug)
        p.setdefault("return_detector_debug", keep_debug)
        p.setdefault("return_noise_psd", keep_debug)
        key = self._params_cache_key(p)
        proc = self._proc_cache.get(key)
        if proc is None:
            proc = SpectralNoiseProcessor(device=self._device, fft_f64=self._fft_f64)
            proc.setup(p)
            self._proc_cache[key] = proc
        return p, proc

    def run(self, audio_data: np.ndarray, params: Dict[str, Any]) -> Tuple[Dict[str, Any], Dict[str, Any]]:
        self._validate_audio(audio_data, params)
        return self.run_batch([audio_data], params, _validated=True)[0]

    def run_batch(self, audio_list: Sequence[np.ndarray], params: Dict[str, Any], _validated: bool = False
                  ) -> List[Tuple[Dict[str, Any], Dict[str, Any]]]:
        if not _validated:
            for a in audio_list:
                self._validate_audio(a, params)
        p, proc = self._prepare(params)
        cfg = proc.cfg
        sr = int(p.get("sample_rate", 11162))
        min_frames = max(1, int(p.get("clip_rain_min_frames", 1)))
        t0 = time.perf_counter()
        outs = proc.process_batch(audio_list, sr=sr, clip_rain_min_frames=min_frames, with_stats=True, with_events=False)
        latency = (time.perf_counter() - t0) / max(1, len(audio_list))
        self.last_host_call_s = getattr(proc, "last_host_call_s", 0.0)    # time inside the C-ABI call of this batch
        keep_features = bool(p.get("keep_state_features", True))
        keep_debug = bool(p.get("keep_state_debug", False))
        keep_spectra = bool(p.get("keep_state_spectra", False))
        keep_audio = bool(p.get("keep_state_audio", False))
        keep_config = bool(p.get("keep_state_config", False))
        name = self.name
        results = []
        for audio, out in zip(audio_list, outs):
            st = out.pop("_clip_stats").tolist()      # [clip id, count, fraction, is_rain, conf, median conf, mean dB, median dB]
            out.pop("_event_idx")
            fc = out["frame_class"]
            count = int(st[1])
            frac = count / fc.size if fc.size else 0.0
            is_rain, conf, median_conf = st[3] > 0.5, st[4], st[5]
            metrics: Dict[str, Any] = {
                "rain_frame_fraction": frac,
                "clip_rain_fraction": frac,
                "rain_frame_count": count,
                "clip_is_rain": is_rain,
                "clip_rain_conf": conf,
                "median_rain_conf": median_conf,
                "clip_rain_min_frames": min_frames,
                "latency_s": latency,
            }
            if out.get("noise_psd") is not None:
                metrics["mean_noise_floor_db"] = st[6]
                metrics["median_noise_floor_db"] = st[7]
            state: Dict[str, Any] = {
                "frame_class": fc, "times": out["times"],
                "rain_conf": out["rain_conf"], "noise_conf": out["noise_conf"],
                "rain_frame_count": count, "clip_rain_fraction": frac,
                "clip_is_rain": is_rain, "clip_rain_conf": conf,
                "median_rain_conf": median_conf, "clip_rain_min_frames": min_frames,
                "latency_s": latency, "processor": name,
            }
            if keep_features:
                state["features"] = out.get("features")
            if keep_debug:
                for k in ("debug", "det_debug", "freqs", "noise_psd"):
                    if k in out:
                        state[k] = out[k]
            if keep_spectra:
                state["S"] = out.get("S")
                state["S_hat"] = out.get("S_hat")
            if keep_audio:
                state["input_audio"] = audio
                if "x_filt" in out:
                    state["filtered_audio"] = out["x_filt"]
                if "y" in out:
                    state["output_audio"] = out["y"]
            if keep_config:
                state["config"] = cfg
            results.append((metrics, state))
        return results
