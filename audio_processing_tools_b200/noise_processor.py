"""NoiseProcessor: framework-level noise-floor processor (reference: noise_processor.py:15-129).

The reference class is stale against its own engine: ``_run_spectral_noise`` reads ``out["y"]``, ``out["is_rain"]`` and
``out["x_hp"]`` (:98-102), which ``SpectralNoiseProcessor.process`` no longer emits, so ``run`` raises ``KeyError``
(SURVEY.md App. B).  This twin implements the contract that code states (:104-127):

  results  mean_noise_floor_db, median_noise_floor_db  -- np.mean / np.median of 10*log10(noise_psd[band] + eps),
           rain_frame_fraction                          -- np.mean(is_rain) with ``is_rain := frame_class == RAIN``,
           latency_s
  state    is_rain, freqs, times, processor, latency_s always; the entries the reference lists beside them
           (noise_psd, debug, S, S_hat, input_audio = prefiltered waveform, denoised_audio, config) behind the same
           ``keep_state_debug`` / ``keep_state_spectra`` / ``keep_state_audio`` / ``keep_state_config`` switches
           RainDetectorProcessor uses (``keep_state_full=True`` turns all four on): a batch of 1 000 ten-minute clips cannot return ~100 MB of
           spectra per clip, and under default flags nothing but the labels and the clip statistics leaves the GPU.

Numbers come from the same kernels as RainDetectorProcessor (the clip statistics rows); the parity oracle is
``RainDetectorProcessor(keep_state_debug=True)`` of the reference (same ``noise_psd``).
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Any, Dict, List, Sequence, Tuple

import numpy as np

from .edge.rain_signal_processor import RainDetectorProcessor
from .processors import BaseProcessor


@dataclass
class NoiseProcessor(BaseProcessor):
    device: int = 0

    def __post_init__(self):
        self._det = None

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_det"] = None
        return d

    def _detector(self) -> RainDetectorProcessor:
        if getattr(self, "_det", None) is None:
            self._det = RainDetectorProcessor(name=self.name, device=self.device)
        return self._det

    def run(self, audio_data: np.ndarray, params: Dict[str, Any]) -> Tuple[Dict[str, Any], Dict[str, Any]]:
        self._validate_audio(audio_data, params)
        return self.run_batch([audio_data], params, _validated=True)[0]

    def run_batch(self, audio_list: Sequence[np.ndarray], params: Dict[str, Any], _validated: bool = False
                  ) -> List[Tuple[Dict[str, Any], Dict[str, Any]]]:
        if not _validated:
            for a in audio_list:
                self._validate_audio(a, params)
        p = dict(params)
        if bool(p.get("keep_state_full", False)):
            for k in ("keep_state_debug", "keep_state_spectra", "keep_state_audio", "keep_state_config"):
                p[k] = True
        keep_debug = bool(p.get("keep_state_debug", False))
        keep_spectra = bool(p.get("keep_state_spectra", False))
        keep_audio = bool(p.get("keep_state_audio", False))
        # the noise-floor statistics need the suppressor's noise PSD pass: both switches that skip it are refused
        if bool(p.get("suppressor_bypass", False)) or bool(p.get("classifier_only_mode", False)):
            raise ValueError("NoiseProcessor needs the noise-PSD pass: suppressor_bypass / classifier_only_mode are not allowed")
        p, proc = self._detector()._prepare(p)
        cfg = proc.cfg
        sr = int(p.get("sample_rate", 11162))
        t0 = time.perf_counter()
        outs = proc.process_batch(audio_list, sr=sr, with_stats=True, with_events=False)
        latency = (time.perf_counter() - t0) / max(1, len(audio_list))
        res = []
        for out in outs:
            stats = out["_clip_stats"]
            fc = out["frame_class"]
            metrics = {"mean_noise_floor_db": float(stats[6]),
                       "median_noise_floor_db": float(stats[7]),
                       "rain_frame_fraction": float(int(stats[1]) / fc.size) if fc.size else 0.0,
                       "latency_s": latency}
            state: Dict[str, Any] = {"is_rain": fc == 2, "freqs": out["freqs"], "times": out["times"],
                                     "processor": self.name, "latency_s": latency}
            if keep_debug:
                state["noise_psd"] = out.get("noise_psd")
                state["debug"] = out.get("debug")
            if keep_spectra:
                state["S"] = out.get("S")
                state["S_hat"] = out.get("S_hat")
            if bool(p.get("keep_state_config", False)):
                state["config"] = cfg
            if keep_audio:
                state["input_audio"] = out.get("x_filt")
                state["denoised_audio"] = out.get("y")
            res.append((metrics, state))
        return res
