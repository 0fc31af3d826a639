"""Deterministic synthetic rain-sensor clips (normative).

Test / benchmark data only -- the reference ships no audio.  A clip is a
Gaussian floor + amplitude-modulated 90 Hz "wind" rumble + Poisson rain drops,
each drop a 30 ms sum of five exponentially decaying sinusoids (one per mode
band), clipped to +-1 and quantised to int16.  The RNG draw order is part of
the spec; the int16 buffer is the canonical input.
"""
from __future__ import annotations

import numpy as np

FS = 11162
MODES = [(450.0, 650.0), (800.0, 1050.0), (1500.0, 1800.0), (2350.0, 2550.0), (3150.0, 3350.0)]
LAMBDAS = (0.0, 0.5, 3.0, 10.0)


def synth_clip_i16(seconds, seed, lam, noise_rms=0.01, fs=FS):
    rng = np.random.default_rng(seed)
    n = int(fs * seconds)
    x = rng.standard_normal(n).astype(np.float64) * noise_rms
    t = np.arange(n) / fs
    x += 0.02 * np.sin(2 * np.pi * 90 * t + rng.uniform(0, 6)) * (0.5 + 0.5 * np.sin(2 * np.pi * 0.2 * t))
    nd = rng.poisson(lam * seconds)
    pos = rng.integers(0, n - 2048, nd)
    L = int(0.03 * fs)
    tt = np.arange(L) / fs
    for p in pos:
        a = rng.uniform(0.05, 0.4)
        pulse = np.zeros(L)
        for (lo, hi), w in zip(MODES, [1.0, 0.6, 0.4, 0.3, 0.2]):
            f = rng.uniform(lo, hi)
            pulse += w * np.sin(2 * np.pi * f * tt + rng.uniform(0, 6)) * np.exp(-tt / 0.006)
        x[p:p + L] += a * pulse
    return np.round(np.clip(x, -1, 1) * 32767).astype(np.int16)


def quiet_clip_i16(seconds, seed, burst_at=(), fs=FS):
    """Near-silent clip (far below the DSD emulator's 0.6 rain-energy threshold) with optional loud
    550 Hz bursts at the given times (seconds): exercises the emulator's not-raining / rain-check branches."""
    rng = np.random.default_rng(seed)
    n = int(fs * seconds)
    x = rng.standard_normal(n) * 0.0002
    for t0 in burst_at:
        i = int(t0 * fs)
        tt = np.arange(int(0.05 * fs)) / fs
        x[i:i + tt.size] += 0.3 * np.sin(2 * np.pi * 550.0 * tt) * np.exp(-tt / 0.01)
    return np.round(np.clip(x, -1, 1) * 32767).astype(np.int16)


def pcm_to_f32(i16):
    """audio_io.safe_to_float semantics (reference audio_io.py:71-72)."""
    return np.asarray(i16, dtype=np.int16).astype(np.float32) / np.float32(32767.0)


def batch_clip_spec(index):
    """(seed, lam) used by benchmark batches: seed = clip index, lam cycles."""
    return int(index), LAMBDAS[int(index) % 4]


def default_params(check_duration=60, **extra):
    p = {"sample_rate": FS, "check_duration": check_duration,
         "detector": {"mode_bands": [tuple(m) for m in MODES]}}
    p.update(extra)
    return p
