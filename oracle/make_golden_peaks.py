#!/usr/bin/env python
"""Golden vectors for the optional peak-structure features (peak_features_enable,
edge/rain_frame_classifier.py:670-683, :761-843): UNMODIFIED reference through the harness.

    python oracle/make_golden_peaks.py
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
from audio_processing_tools_b200.synth import default_params, pcm_to_f32, synth_clip_i16  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
CASES = (("peaks_s16_l10_20s", 20.0, 16, 10.0, {}),
         ("peaks_s17_l3_15s_alt", 15.0, 17, 3.0, {"peak_top_p": 4, "primary_top_m": 2, "peak_prominence_db": 2.0,
                                                   "peak_min_db_above_floor": 4.0, "peak_ratio_min": 0.4,
                                                   "peak_valid_prom_min_db": 2.5, "peak_valid_prom_max_db": 9.0}))


def main():
    for name, seconds, seed, lam, extra in CASES:
        pcm = synth_clip_i16(seconds, seed, lam)
        params = default_params(check_duration=int(seconds), keep_state_debug=True)
        params["detector"].update({"peak_features_enable": True, **extra})
        m, s = RainDetectorProcessor().run(pcm_to_f32(pcm), params)
        dd = s["det_debug"]
        d = {"meta": np.array(json.dumps({"seconds": seconds, "seed": seed, "lam": lam, "detector_extra": extra,
                                          "pcm_sha1": hashlib.sha1(pcm.tobytes()).hexdigest()})),
             "frame_class": np.asarray(s["frame_class"], dtype=np.int8)}
        for k in ("peak_ratio", "peak_gate_score", "peak_valid_count", "peak_count_by_mode"):
            d[k] = np.asarray(dd[k])
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **d)
        print(path, {k: (v.shape, str(v.dtype)) for k, v in d.items() if k != "meta"}, "valid peaks", int(d["peak_valid_count"].sum()),
              "gate pass", int((d["peak_gate_score"] >= 1).sum()))


if __name__ == "__main__":
    main()
