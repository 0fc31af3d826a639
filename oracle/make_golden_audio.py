#!/usr/bin/env python
"""Golden vectors for the suppressed output audio (compute_output_audio: gain -> S_hat -> ISTFT,
edge/rain_signal_processor.py:1113-1128).  Runs the UNMODIFIED reference through the harness; the inverse
STFT itself comes from the harness' librosa stand-in (librosa is not installable offline), so parity at that
call is unpinned in the same sense as at librosa.stft.  Test infrastructure.

    python oracle/make_golden_audio.py
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


def main():
    for name, seconds, seed, lam in (("audio_s14_l3_6s", 6.0, 14, 3.0), ("audio_s15_l10_5s", 5.03, 15, 10.0)):
        pcm = synth_clip_i16(seconds, seed, lam)
        params = default_params(check_duration=int(seconds), keep_state_audio=True, keep_state_spectra=True, keep_state_debug=True)
        m, s = RainDetectorProcessor().run(pcm_to_f32(pcm), params)
        d = {"meta": np.array(json.dumps({"seconds": seconds, "seed": seed, "lam": lam,
                                          "pcm_sha1": hashlib.sha1(pcm.tobytes()).hexdigest()})),
             "output_audio": np.asarray(s["output_audio"], dtype=np.float32),
             "filtered_audio": np.asarray(s["filtered_audio"], dtype=np.float32),
             "frame_class": np.asarray(s["frame_class"], dtype=np.int8)}
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **d)
        print(path, d["output_audio"].shape, os.path.getsize(path) // 1024, "KiB", float(np.abs(d["output_audio"]).max()))


if __name__ == "__main__":
    main()
