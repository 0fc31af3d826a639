"""CPU restatement of the drop-size-distribution emulator.  TEST INFRASTRUCTURE ONLY.

Restates host_analysis/device_dsd_processing_emulator.py of the reference
(`DsdProcessingEmualtor`: __init__ :16-84 constants, set_audio_timestamp :105-114,
process_audio_frame :128-180, calculate_fft_energies :182-202, get_frames_to_next_interval /
get_time_to_next_interval :213-237, process_audio_data :257-314) in two steps: per-frame spectral
quantities for every hop position (vectorised), then the per-minute state machine over them.
Pinned against tests/golden/dsd_cases.npz (outputs of the unmodified reference, oracle/make_golden_dsd.py).
Only tests/ may import this module.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.fft
from scipy.signal import get_window

N_DSD, N_PFT, N_FFT = 32, 30, 38


def constants(fs=11162, frame_length=512):
    dF = fs / frame_length
    c = {"dF": dF, "n_bins": frame_length // 2,
         "rain_lo": int(400 // dF) + 1, "rain_hi": int(700 // dF),
         "pft_lo": int(100 // dF) + 1, "pft_hi": int(1500 // dF) - 1,
         "lwin0": int(300 // dF), "hwin0": int(1000 // dF)}
    c["lwin1"] = c["lwin0"] + N_FFT // 2 - 1
    c["hwin1"] = c["hwin0"] + N_FFT // 2 - 1
    return c


def frame_quantities(x, fs=11162, frame_length=512, hop=512, window=False):
    """drop energy, peak index, peak energy for every frame position i*hop (float64 throughout)."""
    x = np.asarray(x, dtype=np.float64)
    c = constants(fs, frame_length)
    n_frames = 0 if x.size < frame_length else (x.size - frame_length) // hop + 1
    if n_frames == 0:
        return np.zeros(0), np.zeros(0, np.int64), np.zeros(0)
    idx = np.arange(n_frames)[:, None] * hop + np.arange(frame_length)[None, :]
    fr = x[idx]
    if window:
        fr = fr * get_window("hann", frame_length)
    spec = np.abs(scipy.fft.fft(fr, axis=1))
    drop = np.zeros(n_frames)
    for i in range(c["rain_lo"], c["rain_hi"] + 1):      # sequential accumulation like the reference
        drop = drop + spec[:, i]
    pk = np.argmax(spec[:, c["pft_lo"]:c["pft_hi"]], axis=1) + c["pft_lo"]
    return drop, pk, spec[np.arange(n_frames), pk]


def process_audio_data(x, ts, fs=11162, frame_length=512, hop=512, window=False, raining=True):
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    c = constants(fs, frame_length)
    out = []
    if n < frame_length:
        return out
    drop, pk, pke = frame_quantities(x, fs, frame_length, hop, window)
    ts_start = ts - (ts % 60)
    ts_cur = ts
    frame_count = int((ts_cur % 60) * fs / hop)
    pos = 0   # frames consumed (processed or skipped): the reference slices audio_data[hop:]
    energy = np.zeros(N_DSD + N_PFT + N_FFT)
    peak_hist = np.zeros(c["n_bins"])
    freq_hist = np.zeros(c["n_bins"])
    logbase = math.log(1.13)

    def remaining():
        return n - pos * hop

    def tti():
        t = 60 - (ts_cur % 60)
        if t < hop / fs:
            t += 60
        return t

    def do_frame():
        nonlocal pos, frame_count, ts_cur
        i = pos
        if pke[i] != 0:
            peak_hist[pk[i]] += 1
            freq_hist[pk[i]] += pke[i]
        nxt = int(((ts_cur + hop / fs) % 60) / 2)
        cur = int((ts_cur % 60) / 2)
        energy[N_DSD + cur] = np.argmax(peak_hist)
        if nxt != cur:
            peak_hist.fill(0)
        if drop[i] > 0.6:
            h = math.floor(math.log(1 + (drop[i] - 0.6) * 0.6) / logbase)
            energy[min(max(h, 0), N_DSD - 1)] += 1
        pos += 1
        frame_count += 1
        ts_cur = ts_start + frame_count * hop / fs

    for _ in range(math.ceil(n / (fs * 60))):
        energy.fill(0); peak_hist.fill(0); freq_hist.fill(0)
        alive = True
        if raining:
            frames = min(int(tti() * fs / hop), int(remaining() / hop))
            if remaining() < frame_length:
                frames = 0
            for _f in range(frames):
                if remaining() >= frame_length:
                    do_frame()
            for i in range(c["n_bins"]):
                j = min(int(math.log(freq_hist[i] + 2.719) * 25.0), 255)
                if c["lwin0"] <= i <= c["lwin1"]:
                    energy[N_DSD + N_PFT + i - c["lwin0"]] = j
                if c["hwin0"] != c["lwin1"] and c["hwin0"] <= i <= c["hwin1"]:
                    energy[N_DSD + N_PFT + (i - c["hwin0"]) + N_FFT // 2] = j
        else:
            check = ts_cur + tti() - 3
            while ts_cur < check:
                pos += 1
                frame_count += 1
                ts_cur = ts_start + frame_count * hop / fs
                if remaining() < frame_length:
                    alive = False
                    break
            if not alive:
                break
            energy.fill(0); peak_hist.fill(0); freq_hist.fill(0)
            while ts_cur < check + 3:
                if remaining() >= frame_length:
                    do_frame()
                else:
                    alive = False
                    break
            if not alive:
                break
        raining = bool(np.any(energy[:N_DSD] != 0))
        out.append(energy.copy())
        if remaining() < frame_length:
            break
    return out
