#!/usr/bin/env python
"""Golden vectors for the legacy "RoE" rain detector: the UNMODIFIED reference
`rain_detection_algo` (edge/dsp_rain_detection.py:2566-2575) through the harness on synthetic clips, called the
way `processors.RainProcessor.run` calls it (`fn(audio, **params)`, processors.py:112-117).

The reference keeps module-level state between calls (`max_harmonics`, dsp_rain_detection.py:1141,1394-1403):
the cases run in ONE process in the listed order and the value each call started from is recorded.

    python oracle/make_golden_roe.py
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

import audio_processing_tools.edge.dsp_rain_detection as roe  # noqa: E402
from audio_processing_tools_b200.synth import pcm_to_f32, quiet_clip_i16, synth_clip_i16  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden", "roe_cases.npz")
STATE_ARRAYS = ("raining", "kurtosis", "crest_factor", "diff_energy", "energy_list", "min_energy", "times", "Nov0", "novt", "novk")
CASES = [   # name, kind, seconds, seed, lam / bursts, parameter overrides
    ("floor10", "synth", 10, 73, 0.0, {}),
    ("rain10_l3", "synth", 10, 1, 3.0, {}),
    ("rain10_l10", "synth", 10, 72, 10.0, {}),
    ("rain7_short_last_part", "synth", 7.3, 74, 10.0, {}),
    ("quiet10_bursts", "quiet", 10, 75, (2.5, 4.2, 7.0), {}),
    ("rain10_no_fp_fn", "synth", 10, 76, 10.0, {"handle_fp": False, "handle_fn": False}),
    ("rain20_thresholds", "synth", 20, 77, 10.0, {"check_duration": 20, "harmonic_threshold": [3.5, 3.0, 3.0, 3.0, 3.0, 3.0],
                                                  "kurtosis_thr": 2.0, "crest_thr": 3.0, "diff_energy_thr": 4.0, "min_drop_count": 0.2}),
    ("rain10_l30", "synth", 10, 78, 30.0, {}),
    # a base band of 560..860 Hz puts the estimated natural frequency above 550 Hz: `max_harmonics` drops to 5 and stays there
    ("rain10_fn560", "synth", 10, 79, 10.0, {"fn": 560, "n_freq_range": [400, 900]}),
    ("rain10_after_state_change", "synth", 10, 80, 10.0, {}),
]


def make_pcm(kind, seconds, seed, arg):
    return synth_clip_i16(seconds, seed, arg) if kind == "synth" else quiet_clip_i16(seconds, seed, tuple(arg))


def main():
    d, meta = {}, []
    for name, kind, seconds, seed, arg, extra in CASES:
        pcm = make_pcm(kind, seconds, seed, arg)
        params = dict(roe.default_params)
        params.update(extra)
        mh_in = int(roe.max_harmonics)
        drops, frain_mean, st = roe.rain_detection_algo(pcm_to_f32(pcm), **params)
        for k in STATE_ARRAYS:
            if k in st:
                d[f"{name}__{k}"] = np.asarray(st[k])
        if "nov" in st:
            d[f"{name}__nov_rows"] = np.array([len(x) for x in st["nov"]])
        scal = {"rain_drops": int(drops), "frain_mean": float(frain_mean), "rain_drop_count": int(st["rain_drop_count"]),
                "rain_peaks_count": int(st["rain_peaks_count"]), "rain_drop_count_mod": int(st["rain_drop_count_mod"]),
                "max_harmonics_in": mh_in, "max_harmonics_out": int(roe.max_harmonics)}
        d[f"{name}__scalars"] = np.array(json.dumps(scal))
        meta.append({"name": name, "kind": kind, "seconds": seconds, "seed": seed, "arg": arg, "extra": extra,
                     "pcm_sha1": hashlib.sha1(pcm.tobytes()).hexdigest()})
        print(name, scal, "raining frames", int((np.asarray(st["raining"]) >= 1).sum()), flush=True)
    d["meta"] = np.array(json.dumps(meta))
    d["default_params"] = np.array(json.dumps(roe.default_params))
    np.savez_compressed(OUT, **d)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
