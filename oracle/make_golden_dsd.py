#!/usr/bin/env python
"""Golden vectors for the drop-size-distribution emulator (transform.py's compute path).

Runs the UNMODIFIED reference `DsdProcessingEmualtor.process_audio_data`
(host_analysis/device_dsd_processing_emulator.py:274-314) in the build container on synthetic
int16 clips scaled as `parse.pcm_to_float` does (/32768 in float64) and freezes the per-minute
100-vectors (32 drop-size bins + 30 peak-frequency slots + 38 FFT energies).  Test infrastructure.

    python oracle/make_golden_dsd.py
"""
from __future__ import annotations

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

from audio_processing_tools.host_analysis.device_dsd_processing_emulator import DsdProcessingEmualtor  # noqa: E402
from audio_processing_tools_b200.synth import FS, quiet_clip_i16, synth_clip_i16  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden", "dsd_cases.npz")


CASES = [   # name, generator kind, seconds, seed, lam / burst times, ts, window
    ("rain60_l3", "synth", 60, 51, 3.0, 0, False),
    ("rain60_l10", "synth", 60, 52, 10.0, 0, False),
    ("floor60", "synth", 60, 53, 0.0, 0, False),
    ("rain150", "synth", 150, 54, 3.0, 0, False),
    ("rain200_ts37", "synth", 200, 55, 0.5, 37, False),
    ("quiet250", "quiet", 250, 56, (118.5, 238.2), 0, False),
    ("quiet130_ts1000003", "quiet", 130, 57, (55.0,), 1000003, False),
    ("rain60_window", "synth", 60, 58, 3.0, 0, True),
    ("short_0p5min", "synth", 31.7, 59, 3.0, 12, False),
]


def make_pcm(kind, seconds, seed, arg):
    return synth_clip_i16(seconds, seed, arg) if kind == "synth" else quiet_clip_i16(seconds, seed, arg)


def main():
    d = {}
    meta = []
    import hashlib
    for name, kind, seconds, seed, arg, ts, win in CASES:
        pcm = make_pcm(kind, seconds, seed, arg)
        em = DsdProcessingEmualtor(fs=FS, frame_length=512, hop_length=512, bwindow=win, ts=0, verbose=False)
        out = em.process_audio_data(audio_data=pcm.astype(np.float64) / 32768.0, ts=ts)
        arr = np.asarray(out, dtype=np.float64).reshape(len(out), 100)
        d[name + "__out"] = arr
        meta.append({"name": name, "kind": kind, "seconds": seconds, "seed": seed, "arg": arg, "ts": ts, "window": win,
                     "minutes": int(arr.shape[0]), "pcm_sha1": hashlib.sha1(pcm.tobytes()).hexdigest()})
        print(name, arr.shape, "dsd counts", arr[:, :32].sum(axis=1), flush=True)
    d["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(OUT, **d)
    print(OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
