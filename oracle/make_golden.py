#!/usr/bin/env python
"""Generate golden vectors by RUNNING THE UNMODIFIED REFERENCE (test infrastructure).

Runs in the build container only (needs /root/reference or $APT_REFERENCE).
The reference is imported through oracle/refharness (librosa stand-in + import
stubs) and driven through its own public entry point
`RainDetectorProcessor.run` (edge/rain_signal_processor.py:1223) with
keep_state_debug / keep_state_spectra, on the normative synthetic clips
(SURVEY.md Appendix C).  Outputs small .npz fixtures under tests/golden/ that
travel to the GPU box (the reference itself cannot).

    python oracle/make_golden.py            # regenerate everything
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)

import refharness  # noqa: E402

refharness.install()

from audio_processing_tools.edge.rain_signal_processor import RainDetectorProcessor  # noqa: E402
from audio_processing_tools.edge.rain_frame_classifier import (  # noqa: E402
    causal_stochastic_low_quantile_baseline)
from audio_processing_tools.edge.feature_extraction import RAW_SPECTRAL_FEATURE_NAMES  # noqa: E402
from audio_processing_tools_b200.synth import (  # noqa: E402
    FS, MODES, default_params, pcm_to_f32, synth_clip_i16)

OUT = os.path.join(REPO, "tests", "golden")

DET_KEYS = ("mode_flux_score", "mode_flux_score_gated", "primary_mode_flux",
            "support_mode_flux_1", "support_mode_flux_2", "support_mode_flux_3",
            "support_mode_flux_4", "td_crest_factor", "td_kurtosis",
            "td_block_energy_crest", "td_block_peak_width_50",
            "td_block_post_pre_energy_ratio", "td_gate_mask")
METRIC_KEYS = ("rain_frame_fraction", "clip_rain_fraction", "rain_frame_count",
               "clip_is_rain", "clip_rain_conf", "median_rain_conf",
               "clip_rain_min_frames", "mean_noise_floor_db", "median_noise_floor_db")


def run_reference(pcm, seconds, extra=None, spectra=False):
    params = default_params(check_duration=seconds, keep_state_debug=True,
                            keep_state_spectra=spectra)
    if extra:
        for k, v in extra.items():
            if k == "detector":
                params["detector"] = {**params["detector"], **v}
            else:
                params[k] = v
    proc = RainDetectorProcessor()
    metrics, state = proc.run(pcm_to_f32(pcm), params)
    return metrics, state, params


def pack(pcm, seconds, seed, lam, metrics, state, *, level, params):
    fc = np.asarray(state["frame_class"], dtype=np.int8)
    d = {
        "meta": np.array(json.dumps({
            "seconds": seconds, "seed": seed, "lam": lam, "fs": FS,
            "pcm_sha1": hashlib.sha1(pcm.tobytes()).hexdigest(),
            "params": {k: v for k, v in params.items()
                       if k not in ("keep_state_debug", "keep_state_spectra")},
            "numpy": np.__version__,
        }, default=list)),
        "frame_class": fc,
        "rain_conf": np.asarray(state["rain_conf"], dtype=np.float32),
        "noise_conf": np.asarray(state["noise_conf"], dtype=np.float32),
        "times": np.asarray(state["times"], dtype=np.float32),
        "event_idx": np.flatnonzero(fc == 2).astype(np.int32),
    }
    for k in METRIC_KEYS:
        d["metric_" + k] = np.asarray(metrics[k])
    if level >= 1:
        dd = state["det_debug"]
        for k in DET_KEYS:
            d["det_" + k] = np.asarray(dd[k])
    if level >= 2:
        dbg = state["debug"]
        band = np.asarray(dbg["band_mask"])
        d["pcm"] = pcm
        d["band_mask"] = band
        d["freqs"] = np.asarray(state["freqs"], dtype=np.float32)
        d["S"] = np.ascontiguousarray(np.asarray(state["S"]).T)          # (T, F) complex64
        d["noise_psd_band"] = np.ascontiguousarray(np.asarray(state["noise_psd"])[band].T)
        d["detector_noise_psd_band"] = np.ascontiguousarray(
            np.asarray(dbg["detector_noise_psd"])[band].T)
        d["detector_noise_psd_lag_band"] = np.ascontiguousarray(
            np.asarray(dbg["detector_noise_psd_lag"])[band].T)
        d["G_band"] = np.ascontiguousarray(np.asarray(dbg["G"])[band].T)
        d["np_ratio_median_t"] = np.asarray(dbg["np_ratio_median_t"], dtype=np.float32)
        for k in RAW_SPECTRAL_FEATURE_NAMES:
            d["det_" + k] = np.asarray(dd[k])
    return d


def main():
    os.makedirs(OUT, exist_ok=True)
    index = []

    def emit(name, seconds, seed, lam, level, extra=None):
        pcm = synth_clip_i16(seconds, seed, lam)
        metrics, state, params = run_reference(pcm, seconds, extra, spectra=(level >= 2))
        d = pack(pcm, seconds, seed, lam, metrics, state, level=level, params=params)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **d)
        fc = d["frame_class"]
        index.append({"name": name, "seconds": seconds, "seed": seed, "lam": lam,
                      "level": level, "T": int(fc.size), "rain": int((fc == 2).sum()),
                      "uncertain": int((fc == 1).sum()), "noise": int((fc == 0).sum()),
                      "extra": extra or {}})
        print(index[-1], os.path.getsize(path) // 1024, "KiB", flush=True)

    # level 2: every array the reference exports, short clips (PCM included)
    emit("full_s11_l3_8s", 8, 11, 3.0, 2)
    emit("full_s12_l10_8s", 8, 12, 10.0, 2)
    emit("full_s13_l0_6s", 6, 13, 0.0, 2)
    # level 1: events + per-frame detector features, 60 s (SURVEY.md 8(d) check table)
    for seed, lam in ((0, 0.0), (1, 0.5), (2, 3.0), (3, 10.0), (1, 3.0)):
        emit(f"ev_s{seed}_l{lam:g}_60s", 60, seed, lam, 1)
    # level 0: events + clip stats only, long recursions
    emit("ev_s21_l3_300s", 300, 21, 3.0, 0)
    emit("ev_s22_l10_180s", 180, 22, 10.0, 0)
    # non-default parameters (config precedence, thresholds, kurtosis upper gate, weights)
    emit("alt_s31_l3_20s", 20, 31, 3.0, 1, extra={
        "q": 0.3, "win_sec": 0.8, "ema_up": 0.5, "ema_down": 0.9,
        "noise_psd_max_ratio": 0.9, "clip_rain_min_frames": 3,
        "detector": {"td_gate_threshold": 3.0, "td_kurtosis_upper_threshold": 12.0,
                     "mode_weights": [1.0, 0.8, 0.6, 0.5, 0.4],
                     "new_rain_primary_flux_min": 1.5, "new_rain_min_support_count": 1,
                     "mode_flux_norm_q": 30.0, "mode_flux_norm_win_sec": 0.3,
                     "noise_hi": 0.7, "mode_flux_noise_max": 1.0}})
    emit("alt_s32_l3_20s_4modes", 20, 32, 3.0, 1, extra={
        "operating_band": (300.0, 3000.0),
        "detector": {"mode_bands": [tuple(m) for m in MODES[:4]]}})

    # K8 pinned directly: the baseline recursion on a synthetic flux series
    rng = np.random.default_rng(7)
    x = (rng.gamma(2.0, 4.0, 4000) * (rng.uniform(size=4000) < 0.9)).astype(np.float32)
    base, _ = causal_stochastic_low_quantile_baseline(
        x, q_percent=20.0, samples_per_sec=FS / 128.0, win_sec=0.5, min_hist_sec=0.0,
        floor=1.0, dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "baseline_k8.npz"), x=x, baseline=base)

    with open(os.path.join(OUT, "INDEX.json"), "w") as f:
        json.dump(index, f, indent=1)


if __name__ == "__main__":
    main()
