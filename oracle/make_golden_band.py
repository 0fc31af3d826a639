#!/usr/bin/env python
"""Golden vectors for the band noise estimator: the UNMODIFIED reference
`BandNoiseEstimatorProcessor.run` (edge/band_noise_processor.py:82-281) through the harness on synthetic clips.

    python oracle/make_golden_band.py
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

from audio_processing_tools.edge.band_noise_processor import BandNoiseEstimatorProcessor  # noqa: E402
from audio_processing_tools_b200.synth import pcm_to_f32, quiet_clip_i16, synth_clip_i16  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden", "band_noise_cases.npz")
STATE_KEYS = ("M_band", "E_band", "N_E", "N_E_raw", "subE", "N_sub", "rain_submask", "fft_rain_frame", "G_mag", "M_clean",
              "noise_effective_q", "M_band_fft", "E_band_fft", "E_hpf", "times_s")
CASES = [   # name, kind, seconds, seed, lam / bursts, extra params
    ("rain60_l3", "synth", 60, 61, 3.0, {}),
    ("rain45_l10", "synth", 45, 62, 10.0, {}),
    ("floor40", "synth", 40, 63, 0.0, {}),
    ("quiet90", "quiet", 90, 64, (20.5, 40.2, 41.0, 70.3), {}),
    ("rain60_smooth_replenish", "synth", 60, 65, 10.0, {"smooth_N_E": True, "noise_replenish_from_all_subframes": True,
                                                        "noise_buffer_ttl_frames": 40, "det.k_subframes": 3, "W": 20, "W_min": 5}),
    ("rain50_legacy_triggers", "synth", 50, 66, 3.0, {"det.use_D_trigger": True, "det.use_dE_over_Ehpf": True, "det.M_db": 4.0,
                                                      "det.N_db": 2.0, "q": 0.4, "ema_alpha": 0.5, "beta": 1.5, "gain_floor": 0.2}),
]


def make_pcm(kind, seconds, seed, arg):
    return synth_clip_i16(seconds, seed, arg) if kind == "synth" else quiet_clip_i16(seconds, seed, tuple(arg))


def main():
    d, meta = {}, []
    for name, kind, seconds, seed, arg, extra in CASES:
        pcm = make_pcm(kind, seconds, seed, arg)
        params = {"sample_rate": 11162, "check_duration": seconds, **extra}
        res, st = BandNoiseEstimatorProcessor().run(pcm_to_f32(pcm), params)
        for k in STATE_KEYS:
            d[f"{name}__{k}"] = np.asarray(st[k])
        d[f"{name}__results"] = np.array(json.dumps({k: (v if isinstance(v, (int, float, str)) else float(v)) for k, v in res.items()}))
        d[f"{name}__energy_stats"] = np.array(json.dumps(st["energy_stats"]))
        meta.append({"name": name, "kind": kind, "seconds": seconds, "seed": seed, "arg": arg, "extra": extra,
                     "pcm_sha1": hashlib.sha1(pcm.tobytes()).hexdigest()})
        print(name, "frames", res["n_frames"], "fft_rain_frac", res["fft_rain_frac"], "gain_med", res["gain_med"],
              "rain sub frac", float(np.mean(st["rain_submask"])), flush=True)
    d["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(OUT, **d)
    print(OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
