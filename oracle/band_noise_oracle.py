"""CPU restatement of the band noise estimator.  TEST INFRASTRUCTURE ONLY.

Restates the reference's edge/band_noise_estimator.py (NoiseFrameDetector :107-310: fft_rain_from_power
:156-181, time_rain_mask_from_subE :188-274; BandNoiseEstimator :516-986: filters :575-590, ring buffer
:627-680, process_frame :770-986, telemetry :715-768) and the summary of edge/band_noise_processor.py:82-281,
in two steps: per-frame quantities from whole-clip streaming filters (vectorised), then the per-frame state
machine.  float64 configuration only (the reference default).  Pinned against
tests/golden/band_noise_cases.npz (outputs of the unmodified reference, oracle/make_golden_band.py).
Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np
import scipy.signal as spsig

EPS = 1e-12
DEFAULTS = dict(
    fs=11162, frame_len=512, hp_cutoff_hz=350.0, hp_order=4, band_hz=(400.0, 700.0), bpf_order=4, subframe_len=128,
    subhop=128, W=30, W_min=10, noise_buffer_ttl_frames=200, q=0.3, ema_alpha=1.0, beta=1.0, gain_floor=0.10, eps=1e-12,
    ne_attack_alpha_dry=0.15, ne_attack_alpha_wet=0.02, ne_release_alpha=0.25, smooth_N_E=False, learn_during_rain=False,
    force_learn_all=False, noise_replenish_from_all_subframes=False, noise_replenish_q=0.20,
    noise_replenish_only_when_buffer_not_full=True, noise_q_adapt_enable=True, noise_q_replenish_alpha=0.2,
    noise_q_normal_alpha=0.1)
DET_DEFAULTS = dict(
    M_db=6.0, N_db=3.0, primary_hz=(450.0, 650.0),
    rain_bands_hz=((450.0, 650.0), (800.0, 1050.0), (1500.0, 1800.0), (2350.0, 2550.0), (3150.0, 3350.0)),
    k_subframes=2, band_rise_db=6.0, excess_rise_db=3.0, min_Ehpf=1e-10, min_Eband=1e-12, use_dE_over_Ehpf=False,
    dE_over_Ehpf_thr=0.08, use_D_trigger=False, D_db=6.0)


def build_config(params):
    cfg, det = dict(DEFAULTS), dict(DET_DEFAULTS)
    for k, v in params.items():
        if k.startswith("det."):
            if k[4:] in det:
                det[k[4:]] = v
        elif k in cfg:
            cfg[k] = v
    if "sample_rate" in params:
        cfg["fs"] = int(params["sample_rate"])
    elif "fs" in params:
        cfg["fs"] = int(params["fs"])
    return cfg, det


def hz_to_bin(f, fs, n):
    return int(np.clip(np.round(f * n / fs), 0, n // 2))


def filters(cfg):
    nyq = 0.5 * cfg["fs"]
    hpf = None
    if cfg["hp_cutoff_hz"] > 0:
        hpf = spsig.butter(cfg["hp_order"], np.clip(cfg["hp_cutoff_hz"] / nyq, 1e-6, 0.999), btype="highpass", output="sos")
    lo, hi = cfg["band_hz"]
    w1, w2 = np.clip(lo / nyq, 1e-6, 0.999), np.clip(hi / nyq, 1e-6, 0.999)
    if w2 <= w1:
        w2 = min(0.999, w1 + 1e-3)
    bpf = spsig.butter(cfg["bpf_order"], [w1, w2], btype="bandpass", output="sos")
    return hpf, bpf


def frame_quantities(x, cfg, det):
    """Streaming filters over the whole clip (identical to frame-by-frame filtering with carried state), then
    per-frame / per-subframe energies and the FFT sums."""
    N, sub, S = cfg["frame_len"], cfg["subframe_len"], 1 + (cfg["frame_len"] - cfg["subframe_len"]) // cfg["subhop"]
    n_frames = 1 + (x.size - N) // N if x.size >= N else 0
    x = np.asarray(x, dtype=np.float64)[: n_frames * N]
    hpf, bpf = filters(cfg)
    x0 = float(x[0])
    xh = x
    if hpf is not None:
        xh, _ = spsig.sosfilt(hpf, x, zi=spsig.sosfilt_zi(hpf) * x0)
    xb, _ = spsig.sosfilt(bpf, xh, zi=spsig.sosfilt_zi(bpf) * x0)
    fh, fb = xh.reshape(n_frames, N), xb.reshape(n_frames, N)
    q = {"E_hpf": np.array([float(np.sum(r * r)) for r in fh]),
         "subEhpf": np.array([[np.sum(r[s * sub:(s + 1) * sub] ** 2) for s in range(S)] for r in fh]),
         "Eb": np.array([float(np.sum(r * r)) for r in fb]),
         "subE": np.array([[np.sum(r[s * sub:(s + 1) * sub] ** 2) for s in range(S)] for r in fb])}
    X = np.fft.rfft(fh, n=N, axis=1)
    P = X.real * X.real + X.imag * X.imag
    freqs = np.fft.rfftfreq(N, d=1.0 / cfg["fs"])
    bm = (freqs >= cfg["band_hz"][0]) & (freqs <= cfg["band_hz"][1])
    q["Mb_fft"] = np.array([float(np.sum(np.abs(r)[bm])) for r in X])
    q["Eb_fft"] = np.array([float(np.sum(r[bm])) for r in P])

    def band(r, f0, f1):
        b0, b1 = hz_to_bin(f0, cfg["fs"], N), hz_to_bin(f1, cfg["fs"], N)
        b0, b1 = max(0, min(b0, r.size - 1)), max(0, min(b1, r.size - 1))
        return 0.0 if b1 < b0 else float(np.sum(r[b0:b1 + 1]))
    rain = np.zeros(n_frames)
    for f0, f1 in det["rain_bands_hz"]:
        rain = rain + np.array([band(r, f0, f1) for r in P])
    q["rain_sum"] = rain
    q["primary"] = np.array([band(r, *det["primary_hz"]) for r in P])
    return n_frames, S, q


def run(audio, params):
    cfg, det = build_config(params)
    n_frames, S, fq = frame_quantities(np.asarray(audio, dtype=np.float64), cfg, det)
    W = int(cfg["W"])
    buf, valid, bidx = np.zeros(W), np.zeros(W, bool), np.full(W, -1, np.int64)
    wr = count_valid = since = 0
    noise_ema, q_eff, ne_smooth = 0.0, float(cfg["q"]), 0.0
    prev_rain = prev_prim = prev_Eb = prev_Lb = prev_Lh = None
    hold = 0
    Mr, Nr, Dr = 10.0 ** (det["M_db"] / 10.0), 10.0 ** (det["N_db"] / 10.0), 10.0 ** (det["D_db"] / 10.0)
    st = dict(noise_energy_sum=0.0, rain_energy_sum=0.0, total_energy_sum=0.0, noise_frame_count=0, rain_frame_count=0,
              total_frame_count=0, noise_buffer_valid_count=0, noise_buffer_min_valid_count=0,
              noise_buffer_underflow_frame_count=0, frames_since_noise_update=0, noise_learned_subframe_count=0,
              noise_replenish_count=0, noise_effective_q=0.0)
    out = {k: np.zeros(n_frames) for k in ("M_band", "E_band", "N_E", "N_E_raw", "G_mag", "M_clean", "noise_effective_q")}
    out.update(subE=fq["subE"].copy(), N_sub=np.zeros((n_frames, S)), rain_submask=np.zeros((n_frames, S), bool),
               fft_rain_frame=np.zeros(n_frames, bool), M_band_fft=fq["Mb_fft"], E_band_fft=fq["Eb_fft"], E_hpf=fq["E_hpf"],
               times_s=np.arange(n_frames, dtype=np.float64) * cfg["frame_len"] / cfg["fs"])

    def expire(frame_idx):
        nonlocal count_valid
        ttl = int(cfg["noise_buffer_ttl_frames"])
        if ttl <= 0 or count_valid <= 0:
            return
        stale = valid & ((frame_idx - bidx) > ttl)
        if stale.any():
            n = int(stale.sum())
            valid[stale] = False; buf[stale] = 0.0; bidx[stale] = -1
            count_valid = max(0, count_valid - n)

    def push(v, frame_idx):
        nonlocal wr, count_valid
        if not valid[wr]:
            count_valid += 1
        buf[wr], valid[wr], bidx[wr] = v, True, frame_idx
        wr = (wr + 1) % W

    for i in range(n_frames):
        frame_idx = i + 1
        subE, subEh, Eb = fq["subE"][i], fq["subEhpf"][i], float(fq["Eb"][i])
        # FFT-domain decision
        rs, pr = float(fq["rain_sum"][i]), float(fq["primary"][i])
        if prev_rain is None:
            fft_rain = False
        else:
            fft_rain = (rs > (prev_rain + EPS) * Mr) and (pr > (prev_prim + EPS) * Nr)
        prev_rain, prev_prim = rs, pr
        # time-domain subframe mask
        mask = np.zeros(S, bool)
        for s in range(S):
            e = float(max(subE[s], EPS))
            if hold > 0:
                mask[s] = True
                hold -= 1
            trig = False
            eh = float(subEh[s])
            if eh >= det["min_Ehpf"] and e >= det["min_Eband"]:
                Lb, Lh = 10.0 * float(np.log10(e + EPS)), 10.0 * float(np.log10(eh + EPS))
                if prev_Lb is not None and prev_Lh is not None:
                    dLb, dLh = Lb - prev_Lb, Lh - prev_Lh
                    if dLb >= det["band_rise_db"] and (dLb - dLh) >= det["excess_rise_db"]:
                        trig = True
                prev_Lb, prev_Lh = Lb, Lh
            else:
                prev_Lb = prev_Lh = None
            if not trig and det["use_dE_over_Ehpf"] and prev_Eb is not None:
                if max(e - prev_Eb, 0.0) / (float(max(subEh[s], EPS)) + EPS) >= det["dE_over_Ehpf_thr"]:
                    trig = True
            if not trig and det["use_D_trigger"] and prev_Eb is not None and e > (prev_Eb + EPS) * Dr:
                trig = True
            if trig:
                mask[s] = True
                hold = max(hold, max(0, int(det["k_subframes"]) - 1))
            prev_Eb = e
        if fft_rain:
            mask = np.ones(S, bool)
        expire(frame_idx)
        learn = np.ones(S, bool) if (cfg["force_learn_all"] or cfg["learn_during_rain"]) else ~mask
        learned = 0
        for s in range(S):
            if learn[s]:
                push(float(max(subE[s], cfg["eps"])), frame_idx)
                learned += 1
        repl = 0
        if cfg["noise_replenish_from_all_subframes"] and learned == 0 and \
                ((not cfg["noise_replenish_only_when_buffer_not_full"]) or count_valid < W):
            push(float(max(np.quantile(subE, float(cfg["noise_replenish_q"])), cfg["eps"])), frame_idx)
            repl = 1
        st["noise_learned_subframe_count"] += learned
        st["noise_replenish_count"] += repl
        since = 0 if learned + repl > 0 else since + 1
        if cfg["noise_q_adapt_enable"]:
            if repl:
                q_eff = (1.0 - cfg["noise_q_replenish_alpha"]) * q_eff + cfg["noise_q_replenish_alpha"] * cfg["noise_replenish_q"]
            if learned:
                q_eff = (1.0 - cfg["noise_q_normal_alpha"]) * q_eff + cfg["noise_q_normal_alpha"] * cfg["q"]
            q_eff = float(np.clip(q_eff, 1e-6, 1.0 - 1e-6))
        expire(frame_idx)
        if count_valid < int(cfg["W_min"]):
            noise_ema = ne_smooth = 0.0
            nsub = 0.0
        else:
            qv = float(np.quantile(buf[valid], q_eff))
            a = float(cfg["ema_alpha"])
            noise_ema = (1.0 - a) * noise_ema + a * qv
            nsub = noise_ema
        ne_raw = float(S * nsub)
        if cfg["smooth_N_E"]:
            raining = bool(fft_rain) or bool(mask.any())
            up = float(cfg["ne_attack_alpha_wet"] if raining else cfg["ne_attack_alpha_dry"])
            a = up if ne_raw > ne_smooth else float(cfg["ne_release_alpha"])
            ne_smooth = (1.0 - a) * ne_smooth + a * ne_raw
            ne = ne_smooth
        else:
            ne = ne_raw
        rain_e = float(np.sum(subE[mask])) if mask.any() else 0.0
        dry_e = float(np.sum(subE[~mask])) if (~mask).any() else 0.0
        prev_total = st["total_frame_count"]
        st["total_energy_sum"] += max(Eb, 0.0)
        st["rain_energy_sum"] += rain_e
        st["noise_energy_sum"] += min(max(ne, 0.0), max(dry_e, 0.0))
        st["total_frame_count"] += 1
        st["noise_buffer_valid_count"] = count_valid
        st["noise_buffer_min_valid_count"] = count_valid if prev_total == 0 else min(st["noise_buffer_min_valid_count"], count_valid)
        if count_valid < int(cfg["W_min"]):
            st["noise_buffer_underflow_frame_count"] += 1
        st["frames_since_noise_update"] = since
        st["noise_effective_q"] = q_eff
        if mask.any():
            st["rain_frame_count"] += 1
        else:
            st["noise_frame_count"] += 1
        g = float(np.sqrt(np.clip(max(Eb - cfg["beta"] * ne, 0.0) / (Eb + cfg["eps"]), 0.0, 1.0)))
        g = float(np.clip(g, cfg["gain_floor"], 1.0))
        Mb = float(np.sqrt(max(Eb, 0.0)))
        out["M_band"][i], out["E_band"][i], out["N_E"][i], out["N_E_raw"][i] = Mb, Eb, ne, ne_raw
        out["G_mag"][i], out["M_clean"][i], out["noise_effective_q"][i] = g, Mb * g, q_eff
        out["N_sub"][i, :] = nsub
        out["rain_submask"][i], out["fft_rain_frame"][i] = mask, fft_rain
    st["noise_energy_mean"] = st["noise_energy_sum"] / max(1, st["noise_frame_count"])
    st["rain_energy_mean"] = st["rain_energy_sum"] / max(1, st["rain_frame_count"])
    st["total_energy_mean"] = st["total_energy_sum"] / max(1, st["total_frame_count"])
    out["energy_stats"] = st
    return out
