#!/usr/bin/env python
"""Golden vectors for the frame-size / hop sweep (BASELINE config 5), features stage only.

Runs the UNMODIFIED reference (RainDetectorProcessor.run with n_fft / hop overrides and
keep_state_spectra) in the build container and freezes the spectra and band-energy features
it exports.  Test infrastructure; see make_golden.py for the harness.

    python oracle/make_golden_sweep.py
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
from audio_processing_tools.edge.feature_extraction import RAW_SPECTRAL_FEATURE_NAMES  # noqa: E402
from audio_processing_tools_b200.synth import FS, default_params, pcm_to_f32, synth_clip_i16  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
CASES = ((256, 64, 3.0), (512, 256, 4.0), (1024, 256, 4.0), (2048, 1024, 5.0), (4096, 1024, 6.0))


def main():
    for n_fft, hop, seconds in CASES:
        seed, lam = 40 + n_fft // 256, 3.0
        pcm = synth_clip_i16(seconds, seed, lam)
        params = default_params(check_duration=seconds, keep_state_debug=True, keep_state_spectra=True,
                                n_fft=n_fft, hop=hop)
        metrics, state = RainDetectorProcessor().run(pcm_to_f32(pcm), params)
        dd = state["det_debug"]
        d = {"meta": np.array(json.dumps({"seconds": seconds, "seed": seed, "lam": lam, "fs": FS, "n_fft": n_fft,
                                          "hop": hop, "pcm_sha1": hashlib.sha1(pcm.tobytes()).hexdigest(),
                                          "numpy": np.__version__})),
             "pcm": pcm,
             "S": np.ascontiguousarray(np.asarray(state["S"]).T),
             "freqs": np.asarray(state["freqs"], dtype=np.float32),
             "band_mask": np.asarray(state["debug"]["band_mask"])}
        for k in RAW_SPECTRAL_FEATURE_NAMES:
            d["det_" + k] = np.asarray(dd[k])
        path = os.path.join(OUT, f"sweep_nfft{n_fft}_hop{hop}.npz")
        np.savez_compressed(path, **d)
        print(path, d["S"].shape, os.path.getsize(path) // 1024, "KiB", flush=True)


if __name__ == "__main__":
    main()
