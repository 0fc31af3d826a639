"""NoiseProcessor: framework-level noise-floor processor (reference: noise_processor.py:15-129).

The reference class is stale against its own engine (it raises KeyError('y'));
this twin implements the contract its code states (:104-127): noise-floor dB statistics over the
operating band and the rain-frame fraction, with ``is_rain := frame_class == RAIN``.
"""
from __future__ import annotations

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
        if self._det is None:
            self._det = RainDetectorProcessor(name=self.name, device=self.device)
        return self._det

    def run(self, audio_data: np.ndarray, params: Dict[str, Any]) -> Tuple[Dict[str, Any], Dict[str, Any]]:
        self._validate_audio(audio_data, params)
        return self.run_batch([audio_data], params)[0]

    def run_batch(self, audio_list: Sequence[np.ndarray], params: Dict[str, Any]
                  ) -> List[Tuple[Dict[str, Any], Dict[str, Any]]]:
        for a in audio_list:
            self._validate_audio(a, params)
        p = dict(params)
        p["return_noise_psd"] = True
        p.setdefault("keep_state_debug", True)
        outs = self._detector().run_batch(audio_list, p, _validated=True)
        res = []
        for m, s in outs:
            metrics = {"mean_noise_floor_db": m["mean_noise_floor_db"],
                       "median_noise_floor_db": m["median_noise_floor_db"],
                       "rain_frame_fraction": m["rain_frame_fraction"],
                       "latency_s": m["latency_s"]}
            state = {"noise_psd": s.get("noise_psd"), "is_rain": s["frame_class"] == 2,
                     "freqs": s.get("freqs"), "times": s["times"], "debug": s.get("debug"),
                     "processor": self.name, "latency_s": m["latency_s"]}
            res.append((metrics, state))
        return res
