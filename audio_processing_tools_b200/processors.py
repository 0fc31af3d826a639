"""Processor adapters (the plugin boundary of the framework).

Same names, argument meaning and error behaviour as the reference's processors.py:
``BaseProcessor`` (:29-76), ``RainProcessor`` (:84-142), ``has_processor`` (:144-166).
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Any, Callable, Dict, Tuple

import numpy as np


@dataclass
class BaseProcessor:
    """name + input validation + timing helper shared by all processors."""

    name: str

    def _validate_audio(self, audio_data: np.ndarray, params: Dict[str, Any]) -> None:
        if not isinstance(audio_data, np.ndarray):
            raise TypeError(f"audio_data must be a NumPy array, got {type(audio_data)}")
        if audio_data.ndim != 1:
            raise ValueError(f"audio_data must be 1-D, got shape {audio_data.shape}")
        sr, dur = params.get("sample_rate"), params.get("check_duration")
        if sr is not None and dur is not None:
            need = int(sr * dur)
            if audio_data.size < need:
                raise ValueError(f"audio_data too short: {audio_data.size} < required {need} samples")

    def _with_timing(self, func: Callable[..., Any], *args, **kwargs) -> Tuple[Any, float]:
        t0 = time.perf_counter()
        out = func(*args, **kwargs)
        return out, time.perf_counter() - t0


@dataclass
class RainProcessor(BaseProcessor):
    """Adapter over any ``fn(audio, **params) -> (rain_drops, frain_mean, state)``."""

    fn: Callable[..., Tuple[int, float, Dict[str, Any]]]

    def run(self, audio_data: np.ndarray, params: Dict[str, Any]) -> Tuple[Dict[str, Any], Dict[str, Any]]:
        self._validate_audio(audio_data, params)
        (rain_drops, frain_mean, state), latency = self._with_timing(self.fn, audio_data, **params)
        results: Dict[str, Any] = {"rain_drops": rain_drops, "frain_mean": frain_mean, "latency_s": latency}
        if isinstance(state, dict):
            for key in ("rain_drop_count", "rain_peaks_count", "rain_drop_count_mod"):
                if key in state:
                    results[key] = state[key]
        state_out: Dict[str, Any] = dict(state) if isinstance(state, dict) else {"state": state}
        state_out["processor"] = self.name
        state_out["latency_s"] = latency
        return results, state_out


    def run_batch(self, audio_list, params: Dict[str, Any]):
        """Additive batch hook of the framework twin: when the wrapped function carries a `batch` companion
        (`fn.batch(list_of_audio, **params) -> list of fn results`, as the GPU RoE detector does) the whole list is one
        call; otherwise the files go through `run` one by one.  Same results / state as `run` per file; `latency_s` is
        the batch time divided by the number of files."""
        batch_fn = getattr(self.fn, "batch", None)
        if batch_fn is None:
            return [self.run(a, params) for a in audio_list]
        for a in audio_list:
            self._validate_audio(a, params)
        outs, latency = self._with_timing(batch_fn, list(audio_list), **params)
        per_file = latency / max(1, len(audio_list))
        packed = []
        for rain_drops, frain_mean, state in outs:
            results: Dict[str, Any] = {"rain_drops": rain_drops, "frain_mean": frain_mean, "latency_s": per_file}
            if isinstance(state, dict):
                for key in ("rain_drop_count", "rain_peaks_count", "rain_drop_count_mod"):
                    if key in state:
                        results[key] = state[key]
            state_out: Dict[str, Any] = dict(state) if isinstance(state, dict) else {"state": state}
            state_out["processor"] = self.name
            state_out["latency_s"] = per_file
            packed.append((results, state_out))
        return packed


def has_processor(processors, name: str) -> bool:
    return any(p.name == name for p in processors)
