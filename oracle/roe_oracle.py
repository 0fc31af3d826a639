"""CPU restatement of the legacy "RoE" rain detector -- TEST INFRASTRUCTURE ONLY.

Follows edge/dsp_rain_detection.py of the reference: `rain_detection_algo` (:2566-2575) ->
`analyse_raw_audio_wrapper` (:2677-2731) -> `analyse_raw_audio_in_parts` (:2603-2636, 2-second parts) ->
`analyse_raw_audio` (:2230-2562), with `calculate_pulse_characteristics` (:657-767), `compute_novelty_spectrum_new`
(:1924-1955), `calculate_snr` / `compute_local_average` (:1892-1922), `find_peaks_in_frequency_range` (:1649-1698),
`bp_filter_frequencies` (:1828-1846), `update_search_freq_range` (:1393-1405), `time_domain_raining_status`
(:770-801) and `combine_raining_status` (:2638-2674).  The reference's module globals become an explicit
configuration; the one global that survives between calls (`max_harmonics`, :1141, :1403) is an explicit argument
and return value.  numpy / scipy calls are the ones the reference makes (sosfilt, rfft, find_peaks, kurtosis).

Pinned against outputs of the unmodified reference: tests/golden/roe_cases.npz (oracle/make_golden_roe.py).
Only tests/ may import this module.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.fft
import scipy.signal as sps
from scipy.stats import kurtosis

FS_ANALYSIS = 11162           # analyse_raw_audio's own `sample_rate` default (:2236); the wrapper never passes another
MAX_DURATION_FW = 2           # :2601

DEFAULT_PARAMS = {            # default_params, :1097-1123
    "sample_rate": 11162, "freq_resolution": 45, "time_resolution_ms": 10, "check_duration": 10,
    "op_freq_range": [400, 3500], "n_freq_range": [400, 700], "fn": 400, "num_harmonics": 6,
    "harmonic_threshold": [4.5, 4.0, 3.5, 3.5, 3.5, 3.5], "max_peaks": 3, "log_factor": 0, "ns_duration_ms": 470, "nf": 0,
    "min_drop_count": 0.3, "rain_drop_min_thr": 3, "rain_drop_max_thr": 50, "rain_peaks_min_thr": 9, "rain_peaks_max_thr": 30,
    "kurtosis_thr": 2.5, "crest_thr": 3.75, "diff_energy_thr": 6.5, "t_band": [400, 3500], "handle_fp": True, "handle_fn": True,
    "enable_nov_wind_dection": False, "enable_energy_peak_detection": False,
}


class RoeConfig:
    """configure_parameters (:1298-1391) without the globals."""

    def __init__(self, sample_rate=11162, freq_resolution=45, time_resolution_ms=10, check_duration=10, op_freq_range=(400, 3500),
                 n_freq_range=(400, 700), fn=400, num_harmonics=6, harmonic_threshold=(4.5, 4.0, 3.5, 3.5, 3.5, 3.5), max_peaks=3,
                 log_factor=0, ns_duration_ms=470, nf=0, min_drop_count=0.3, kurtosis_thr=2.5, crest_thr=3.75, diff_energy_thr=6.5,
                 rain_drop_min_thr=3, rain_drop_max_thr=50, rain_peaks_min_thr=9, rain_peaks_max_thr=30, t_band=(400, 3500),
                 handle_fp=True, handle_fn=True, enable_nov_wind_dection=False, enable_energy_peak_detection=False):
        self.frame_length = 2 ** math.ceil(math.log2(sample_rate / freq_resolution))
        self.hop_length = 2 ** math.ceil(math.log2((time_resolution_ms * sample_rate) / 1000))
        self.check_duration = check_duration
        self.F_natural = fn
        self.op = list(op_freq_range)
        self.natural = list(n_freq_range)
        self.log_factor = log_factor
        self.M = math.ceil(((ns_duration_ms * sample_rate / 1000) / self.hop_length - 1) / 2)
        self.rain_thr = list(harmonic_threshold)
        self.rain_thr_hn = self.rain_thr[0] + self.rain_thr[1] + self.rain_thr[2]
        self.min_drop_count = min_drop_count
        self.max_peaks = max_peaks
        self.nf = nf
        self.process_fp, self.process_fn = handle_fp, handle_fn
        self.search = [list(op_freq_range)] + [[fn * k - 200, fn * k + 300] for k in (2, 3, 4, 5)] + [[fn * 6 - 200, self.op[1]]]


def _stft_mag(y, N, H):
    """|librosa.stft(y, n_fft=N, hop_length=H, window='hann')|, center=True with zero padding (librosa 0.11)."""
    w = sps.get_window("hann", N, fftbins=True)
    yp = np.pad(y, (N // 2, N // 2))
    T = 1 + (yp.size - N) // H
    fr = np.lib.stride_tricks.as_strided(yp, shape=(N, T), strides=(yp.strides[0], H * yp.strides[0]), writeable=False)
    return np.asfortranarray(np.abs(scipy.fft.rfft(w.reshape(-1, 1) * fr, axis=0)))


def _pulse_characteristics(y, T, N, H):
    """calculate_pulse_characteristics (:657-767): kurtosis, crest factor, energy rise per frame (+ one trailing zero)."""
    pad = np.concatenate((np.zeros(H), y, np.zeros(H)))
    sos = sps.butter(4, [400 / (0.5 * FS_ANALYSIS), 900 / (0.5 * FS_ANALYSIS)], btype="band", output="sos")
    filt = sps.sosfilt(sos, pad)
    nfe = 1 + (filt.size - N) // H
    fr = np.lib.stride_tricks.as_strided(filt, shape=(nfe, N), strides=(H * filt.strides[0], filt.strides[0]))
    energy = np.sum(fr ** 2, axis=1)
    k_list, crest, diff, emin = np.zeros(T), np.zeros(T), np.zeros(T), np.zeros(T)
    for i in range(T):
        s, e = i * H, i * H + N
        if e > pad.size:
            break
        x = pad[s:e]
        lo, hi = max(1, i - 30), min(energy.size - 1, i + 31)
        emin[i] = np.min(energy[lo:hi]) if lo < hi else 0
        if i >= 2:
            last = energy[i - 2] if energy[i - 2] < energy[i - 1] else energy[i - 1]
            diff[i] = energy[i] / (last + 1e-12) if energy[i] > last else 0
        if i > 0:
            k_list[i] = kurtosis(x, fisher=True)
            crest[i] = np.max(np.abs(x)) / (np.sqrt(np.mean(x ** 2)) + 1e-12)
    z = [0]
    return {"times": np.concatenate((z, np.arange(T) * H / FS_ANALYSIS)), "kurtosis": np.concatenate((k_list, z)),
            "crest_factor": np.concatenate((crest, z)), "diff_energy": np.concatenate((diff, z)),
            "energy_list": np.concatenate((energy, z)), "min_energy": np.concatenate((emin, z))}


def _novelty(Y, band, N, M, threshold):
    """bp_filter_frequencies + compute_novelty_spectrum_new (:1828-1846, :1924-1955)."""
    f_res = FS_ANALYSIS / N
    i1, i2 = int(band[0] // f_res + 1), int(band[1] // f_res)
    Y1 = Y.copy(order="F")
    Y1[0:i1] = 0
    Y1[i2 + 1:] = 0
    d = np.diff(Y1, n=1, axis=0)
    d[d <= 0] = 0
    nov = np.concatenate((np.sum(d, axis=0), np.array([0])))
    # calculate_snr / compute_local_average: mean of the (at most M // 6, at least 3) smallest values of a +-M window
    L = nov.size
    la = np.zeros(L)
    for m in range(L):
        xd = sorted(nov[max(m - M, 0):min(m + M + 1, L)])
        wl = len(xd)
        if wl > M // 6:
            wl = M // 6
        if wl < 3:
            wl = 3
        la[m] = (1 / wl) * np.sum(xd[:wl])
    la[la <= 0] = np.max(nov) / 5
    nov[nov == 0] = 1
    la[la == 0] = 1
    snr = np.divide(nov, la)
    peaks, _ = sps.find_peaks(snr, prominence=(None, None))
    mask = np.zeros(L)
    mask[peaks] = 1
    unthr = snr * mask
    thr = np.where(snr > threshold, np.minimum(snr, threshold * 1.5), 0.0)
    return thr * mask, unthr


def _peaks_in_range(mag, search, fpeak_range, num_peaks):
    """find_peaks_in_frequency_range (:1649-1698): per frame, the first of the `num_peaks` lowest local maxima inside
    the search band whose frequency lies strictly inside fpeak_range."""
    fn = FS_ANALYSIS / 2
    b1, b2 = int((search[0] * mag.shape[0]) / fn), int((search[1] * mag.shape[0]) / fn)
    rng = mag[b1:b2, :]
    found, fpk = [], []
    for t in range(rng.shape[1]):
        idx, _ = sps.find_peaks(rng[:, t])
        idx = idx + b1
        freqs = (idx * fn) / mag.shape[0]
        f, ok = 0, 0
        for k in range(min(len(idx), num_peaks)):
            if fpeak_range[0] < freqs[k] < fpeak_range[1]:
                ok, f = 1, freqs[k]
                break
        found.append(ok)
        fpk.append(f)
    return found, fpk


def _analyse_part(x, cfg, max_harmonics):
    """analyse_raw_audio (:2230-2562) on one part that is at least one second long."""
    N, H = cfg.frame_length, cfg.hop_length
    nyq = 0.5 * FS_ANALYSIS
    y = sps.sosfilt(sps.butter(8, [cfg.op[0] / nyq, cfg.op[1] / nyq], btype="bandpass", output="sos"), x)
    mag = _stft_mag(y, N, H)
    st = {}
    if cfg.process_fp or cfg.process_fn:
        st.update(_pulse_characteristics(y, mag.shape[1], N, H))
    if cfg.nf != 0:
        raise NameError("estimate_noise_lpf is not defined in the reference (dsp_rain_detection.py:2318)")
    Y = mag if cfg.log_factor == 0 else np.log(1 + cfg.log_factor * mag)
    base_band = [cfg.F_natural, cfg.F_natural + 300]
    novk, novt = _novelty(Y, base_band, N, cfg.M, cfg.rain_thr[0])
    found, fpk = _peaks_in_range(mag, cfg.search[0], base_band, cfg.max_peaks)
    for k in range(len(fpk)):
        if novk[k] != 0 and found[k] == 0:
            novk[k] = 0
            novt[k] = 0
    nz = [f for f in fpk if f != 0]
    frain_mean = np.mean(nz) if nz else 0
    nov, nov1 = [novk.copy()], [novt.copy()]
    # update_search_freq_range (:1393-1405)
    for i in range(1, 6):
        lo = frain_mean * (i + 1) - 200
        if lo < cfg.op[0]:
            lo = cfg.op[0]
        hi = frain_mean * (i + 1) + 300
        if hi > cfg.op[1] + 100:
            max_harmonics = i
        if hi > cfg.op[1]:
            hi = cfg.op[1]
        cfg.search[i] = [lo, hi]
    if cfg.natural[0] <= frain_mean <= cfg.natural[1]:
        for hn in range(1, max_harmonics):
            f1 = frain_mean * (hn + 1) - 100
            band = [f1, f1 + 300]
            novx, novt_h = _novelty(Y, band, N, cfg.M, cfg.rain_thr[hn])
            _, fpk_h = _peaks_in_range(mag, cfg.search[hn], band, cfg.max_peaks)
            for k in range(len(fpk_h)):
                if novx[k] != 0 and fpk_h[k] == 0:
                    novx[k] = 0
            nov.append(novx.copy())
            nov1.append(novt_h.copy())
    for k in range(len(nov[0])):
        if nov[0][k] == 0:
            for j in range(1, len(nov)):
                nov[j][k] = 0
    nov_hn = np.sum(nov, axis=0)
    nov_hn[nov_hn > cfg.rain_thr_hn] = cfg.rain_thr_hn
    nov_hn[nov_hn < cfg.rain_thr_hn] = 0
    st.update({"raining": nov_hn, "Nov0": nov[0], "novt": novt, "novk": novk})
    return int((nov_hn >= 1).sum()), frain_mean, st, max_harmonics


def rain_detection_algo(audio, max_harmonics=6, **params):
    """rain_detection_algo (:2566-2575).  Returns (rain_drops, frain_mean, state, max_harmonics_out)."""
    cfg = RoeConfig(**params)
    duration, offset, count, raining = cfg.check_duration, 0, 0, False
    thr = math.ceil(cfg.min_drop_count * duration)
    merged, frain_mean = {}, 0.0
    n_frames_per_s = FS_ANALYSIS / cfg.frame_length
    while duration > 0:
        part = min(duration, MAX_DURATION_FW)
        x = np.asarray(audio)[int(FS_ANALYSIS * offset):int(FS_ANALYSIS * offset) + int(cfg.frame_length * (part * n_frames_per_s))]
        if len(x) < FS_ANALYSIS:
            drops, frain_mean = 0, 0
        else:
            drops, frain_mean, st, max_harmonics = _analyse_part(x, cfg, max_harmonics)
            for k, v in st.items():
                merged[k] = np.concatenate((merged[k], v)) if k in merged else v
        duration -= part
        offset += part
        count += drops
        if count > thr:
            raining = True
    state = merged
    mod = count
    if cfg.process_fp or cfg.process_fn:
        pk = (state["kurtosis"] > params["kurtosis_thr"]) & (state["crest_factor"] > params["crest_thr"]) & (state["diff_energy"] > params["diff_energy_thr"])
        state["rain_peaks"] = pk
        npk = int((pk > 0).sum())
        # combine_raining_status (:2638-2674)
        if params["handle_fn"] and not raining:
            if count > params["rain_drop_max_thr"] or npk > params["rain_peaks_max_thr"]:
                raining, mod = True, max(count, npk)
        if params["handle_fp"] and raining:
            if npk < params["rain_peaks_min_thr"] or count < thr:
                raining, mod = False, 0
        state.update(rain_drop_count=count, rain_peaks_count=npk, rain_drop_count_mod=mod)
    else:
        state.update(rain_drop_count=count, rain_peaks_count=count, rain_drop_count_mod=count)
    return (mod if raining else 0), frain_mean, state, max_harmonics
