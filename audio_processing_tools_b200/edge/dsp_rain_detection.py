"""GPU twin of the legacy "RoE" rain detector (reference: edge/dsp_rain_detection.py).

`rain_detection_algo(audio_data, **params) -> (rain_drops, frain_mean, state)` has the reference's signature
(:2566-2575), so it is the `fn` that `processors.RainProcessor` wraps (processors.py:84-142); the keyword set is
`configure_parameters`' (:1298-1324; an unknown keyword is a TypeError as there), and `default_params` is the
reference's table (:1097-1123).  `rain_detection_algo_batch` runs a list of clips in one GPU pass.  The compute --
per 2-second part: two band-pass filters, STFT magnitudes, kurtosis / crest / energy rise, band-limited novelty,
local-average SNR, peak masks, the harmonic search -- runs in CUDA behind `apt_roe_run`; there is no CPU path.

Like the reference, the module keeps `max_harmonics` between calls (:1141, :1394-1403: a part whose estimated
natural frequency exceeds 550 Hz lowers it from 6 to 5 for every later part, clip and call of the process).

State keys: raining, kurtosis, crest_factor, diff_energy, energy_list, min_energy, times, Nov0, novt, novk,
rain_peaks, rain_drop_count, rain_peaks_count, rain_drop_count_mod.  Not reproduced: the spectrogram dumps
(spectrum_db*, audio_data, filtered), the per-harmonic lists (nov, nov1) and rain_status_new.  Refused: nf != 0
(the reference calls an undefined function there, :2318), log_factor != 0, the wind / energy-peak experiments,
frame sizes other than 256 / 128.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Any, Dict, List, Sequence, Tuple

import numpy as np
import scipy.signal as spsig

from .. import _lib
from .._staging import upload_clips
from ..engine import AptError, _torch

default_params = {
    "sample_rate": 11162, "freq_resolution": 45, "time_resolution_ms": 10, "check_duration": 10,
    "op_freq_range": [400, 3500], "n_freq_range": [400, 700], "fn": 400, "num_harmonics": 6,
    "harmonic_threshold": [4.5, 4.0, 3.5, 3.5, 3.5, 3.5], "max_peaks": 3, "log_factor": 0, "ns_duration_ms": 470, "nf": 0,
    "min_drop_count": 0.3, "rain_drop_min_thr": 3, "rain_drop_max_thr": 50, "rain_peaks_min_thr": 9, "rain_peaks_max_thr": 30,
    "kurtosis_thr": 2.5, "crest_thr": 3.75, "diff_energy_thr": 6.5, "t_band": [400, 3500], "handle_fp": True, "handle_fn": True,
    "enable_nov_wind_dection": False, "enable_energy_peak_detection": False,
}

FS_ANALYSIS = 11162        # analyse_raw_audio's own sample_rate default (:2236): the wrapper never passes another
MAX_DURATION_FW = 2        # :2601
max_harmonics = default_params["num_harmonics"]     # module state, as in the reference (:1141)

_ctx = None
_device = 0


def configure_parameters(sample_rate=11162, freq_resolution=45, time_resolution_ms=10, check_duration=10, op_freq_range=[400, 3500],
                         n_freq_range=[400, 700], fn=400, num_harmonics=6, harmonic_threshold=[4.5, 4.0, 3.5, 3.5, 3.5, 3.5],
                         max_peaks=3, log_factor=0, ns_duration_ms=470, nf=0, min_drop_count=0.3, kurtosis_thr=2.5, crest_thr=3.75,
                         diff_energy_thr=6.5, rain_drop_min_thr=3, rain_drop_max_thr=50, rain_peaks_min_thr=9, rain_peaks_max_thr=30,
                         t_band=[400, 3500], handle_fp=True, handle_fn=True, enable_nov_wind_dection=False,
                         enable_energy_peak_detection=False) -> Dict[str, Any]:
    """configure_parameters (:1298-1391): the derived quantities the reference keeps in module globals."""
    frame_length = 2 ** math.ceil(math.log2(sample_rate / freq_resolution))
    hop_length = 2 ** math.ceil(math.log2((time_resolution_ms * sample_rate) / 1000))
    if frame_length != 256 or hop_length != 128:
        raise NotImplementedError(f"frame_length={frame_length}, hop_length={hop_length}: the CUDA path is built for 256 / 128")
    if nf != 0:
        raise NameError("name 'estimate_noise_lpf' is not defined")      # what the reference does (:2318)
    if log_factor != 0:
        raise NotImplementedError("log_factor != 0 is not implemented on the CUDA path")
    if enable_nov_wind_dection or enable_energy_peak_detection:
        raise NotImplementedError("the wind / energy-peak experiments are not implemented on the CUDA path")
    thr = list(harmonic_threshold)
    return {
        "frame_length": frame_length, "hop_length": hop_length, "check_duration": check_duration, "F_natural": fn,
        "op": list(op_freq_range), "natural": list(n_freq_range),
        "M": math.ceil(((ns_duration_ms * sample_rate / 1000) / hop_length - 1) / 2),
        "rain_thr": thr, "rain_thr_hn": thr[0] + thr[1] + thr[2], "min_drop_count": min_drop_count, "max_peaks": max_peaks,
        "process_fp": handle_fp, "process_fn": handle_fn,
    }


def _resolve(cfg: Dict[str, Any], params: Dict[str, Any]) -> "_lib.AptRoeParams":
    P = _lib.AptRoeParams()
    P.n_fft, P.hop, P.M = cfg["frame_length"], cfg["hop_length"], cfg["M"]
    if P.M < 2:
        raise NotImplementedError("ns_duration_ms too short for the CUDA path")
    P.wl = max(3, P.M // 6)
    if P.wl > 8:
        raise NotImplementedError("ns_duration_ms too long for the CUDA path (more than 8 values in the local average)")
    if len(cfg["rain_thr"]) < 6:
        raise IndexError("list index out of range")          # rain_thr[hn] in the harmonic loop (:2464)
    P.max_peaks = int(cfg["max_peaks"])
    P.want_td = int(bool(cfg["process_fp"] or cfg["process_fn"]))
    nyq = 0.5 * FS_ANALYSIS
    sos_in = spsig.butter(8, [cfg["op"][0] / nyq, cfg["op"][1] / nyq], btype="bandpass", output="sos")
    sos_td = spsig.butter(4, [400 / nyq, 900 / nyq], btype="band", output="sos")
    P.ns_in, P.ns_td = sos_in.shape[0], sos_td.shape[0]
    for dst, sos in ((P.sos_in, sos_in), (P.sos_td, sos_td)):
        for s in range(sos.shape[0]):
            for j in range(6):
                dst[s][j] = float(sos[s, j])
    w = spsig.get_window("hann", 256, fftbins=True)
    for i in range(256):
        P.window[i] = float(w[i])
    P.fs = float(FS_ANALYSIS)
    P.f_natural, P.op_lo, P.op_hi = float(cfg["F_natural"]), float(cfg["op"][0]), float(cfg["op"][1])
    P.nat_lo, P.nat_hi = float(cfg["natural"][0]), float(cfg["natural"][1])
    P.search0_lo, P.search0_hi = P.op_lo, P.op_hi
    for i in range(6):
        P.rain_thr[i] = float(cfg["rain_thr"][i])
    P.rain_thr_hn = float(cfg["rain_thr_hn"])
    P.rain_drop_threshold = int(math.ceil(cfg["min_drop_count"] * cfg["check_duration"]))
    if P.want_td:           # the reference reads these from the call's keyword dict (:770-801, :2638-2674): KeyError if absent
        P.kurtosis_thr, P.crest_thr, P.diff_energy_thr = float(params["kurtosis_thr"]), float(params["crest_thr"]), float(params["diff_energy_thr"])
        P.handle_fp, P.handle_fn = int(bool(params["handle_fp"])), int(bool(params["handle_fn"]))
        P.rain_peaks_min_thr, P.rain_peaks_max_thr = int(params["rain_peaks_min_thr"]), int(params["rain_peaks_max_thr"])
        P.rain_drop_max_thr = int(params["rain_drop_max_thr"])
        _ = params["rain_drop_min_thr"]
    return P


def _context():
    global _ctx
    if _ctx is None:
        _torch()
        L = _lib.load()
        ctx = C.c_void_p()
        if L.apt_init(_device, C.byref(ctx)) != 0:
            raise AptError("apt_init failed: is a B200 visible?")
        _ctx = ctx
    return _lib.load(), _ctx


def _part_table(lens: Sequence[int], cfg: Dict[str, Any]):
    """analyse_raw_audio_in_parts (:2603-2636): 2-second parts over check_duration; parts shorter than one second
    return early in the reference (:2257-2258) and are not listed.  Returns the table and, per clip, whether the last
    part of the walk was analysed (its frain_mean is the one the call returns)."""
    N = cfg["frame_length"]
    clip, start, plen, last_ok = [], [], [], []
    base = 0
    for c, n in enumerate(lens):
        duration, offset, ok = cfg["check_duration"], 0, False
        while duration > 0:
            part = min(duration, MAX_DURATION_FW)
            a = int(FS_ANALYSIS * offset)
            size = int(N * (part * FS_ANALYSIS / N))
            ln = max(0, min(n, a + size) - min(n, a))
            ok = ln >= FS_ANALYSIS
            if ok:
                clip.append(c); start.append(base + a); plen.append(ln)
            duration -= part
            offset += part
        last_ok.append(ok)
        base += n
    return np.asarray(clip, np.int32), np.asarray(start, np.int64), np.asarray(plen, np.int32), last_ok


def rain_detection_algo_batch(audio_list: Sequence[np.ndarray], **kwargs) -> List[Tuple[int, float, Dict[str, Any]]]:
    global max_harmonics
    cfg = configure_parameters(**kwargs)
    P = _resolve(cfg, kwargs)
    clips = []
    for a in audio_list:
        x = np.asarray(a)
        if x.ndim != 1:
            raise ValueError("audio_data must be a mono ndarray")
        clips.append(x if x.dtype == np.int16 else np.asarray(x, dtype=np.float32))
    is_f32 = clips[0].dtype != np.int16
    if any((c.dtype != np.int16) != is_f32 for c in clips):
        raise TypeError("a batch must be all int16 or all float")
    lens = [c.size for c in clips]
    pclip, pstart, plen, last_ok = _part_table(lens, cfg)
    if P.want_td and (np.bincount(pclip, minlength=len(clips)) == 0).any():
        raise KeyError("raining")              # time_domain_raining_status on an empty state (:787)
    torch = _torch()
    L, ctx = _context()
    dev = torch.device("cuda", _device)
    n_parts = int(pclip.size)
    T1 = plen // 128 + 2                                # frame slots per part
    fo = np.concatenate(([0], np.cumsum(T1))).astype(np.int64)
    d_pcm = upload_clips(torch, dev, clips, np.asarray(lens, dtype=np.int64))      # pinned staging, one asynchronous copy
    d_f = torch.zeros((max(int(fo[-1]), 1), _lib.ROE_FRAME_F), dtype=torch.float64, device=dev)
    d_p = torch.zeros((max(n_parts, 1), _lib.ROE_PART_F), dtype=torch.float64, device=dev)
    d_c = torch.zeros((len(clips), _lib.ROE_CLIP_F), dtype=torch.float64, device=dev)
    mh_out = C.c_int(0)
    with _lib.device_timer(torch, "roe", _device):
        rc = L.apt_roe_run(ctx, C.byref(P), len(clips), d_pcm.data_ptr(), int(is_f32), n_parts,
                           pclip.ctypes.data_as(C.POINTER(C.c_int32)), pstart.ctypes.data_as(C.POINTER(C.c_int64)),
                           plen.ctypes.data_as(C.POINTER(C.c_int32)), int(max_harmonics), d_f.data_ptr(), d_p.data_ptr(), d_c.data_ptr(),
                           C.byref(mh_out), torch.cuda.current_stream(_device).cuda_stream)
    if rc != 0:
        raise AptError(f"apt_roe_run failed ({rc}): {L.apt_last_error(ctx).decode()}")
    max_harmonics = int(mh_out.value)
    # per-frame columns leave the device column-major: a clip's series is one contiguous run of the batch array
    Ft, Pp, Cc = d_f.t().contiguous().cpu().numpy(), d_p.cpu().numpy(), d_c.cpu().numpy()
    # parts of a clip are consecutive (pclip ascending): [p0[c], p1[c])
    p0 = np.searchsorted(pclip, np.arange(len(clips)), side="left")
    p1 = np.searchsorted(pclip, np.arange(len(clips)), side="right")
    kurt_thr, crest_thr, de_thr = P.kurtosis_thr, P.crest_thr, P.diff_energy_thr
    want_td = bool(P.want_td)
    time_rows = {}           # frame times of a part by its slot count (shared by every part of that length)
    outs = []
    for c in range(len(clips)):
        a, b = int(p0[c]), int(p1[c])
        f0, f1 = (int(fo[a]), int(fo[b])) if b > a else (0, 0)
        blk = Ft[:8, f0:f1].copy()                     # one block per clip, a row per series
        state: Dict[str, Any] = {"raining": blk[0], "Nov0": blk[6], "novk": blk[6].copy(), "novt": blk[7]}
        if want_td:
            tl = []
            for q in range(a, b):
                t1 = int(T1[q])
                row = time_rows.get(t1)
                if row is None:
                    row = time_rows[t1] = np.concatenate(([0.0], np.arange(t1 - 1) * 128 / FS_ANALYSIS))
                tl.append(row)
            state.update(kurtosis=blk[1], crest_factor=blk[2], diff_energy=blk[3], energy_list=blk[4], min_energy=blk[5],
                         times=np.concatenate(tl))
            state["rain_peaks"] = (blk[1] > kurt_thr) & (blk[2] > crest_thr) & (blk[3] > de_thr)
        state.update(rain_drop_count=int(Cc[c, 1]), rain_peaks_count=int(Cc[c, 2]), rain_drop_count_mod=int(Cc[c, 3]))
        frain_mean = float(Pp[b - 1, 0]) if (b > a and last_ok[c]) else 0
        outs.append((int(Cc[c, 0]), frain_mean, state))
    return outs


def rain_detection_algo(audio_data, **kwargs):
    """rain_detection_algo (:2566-2575)."""
    return rain_detection_algo_batch([audio_data], **kwargs)[0]


rain_detection_algo.batch = rain_detection_algo_batch      # RainProcessor.run_batch picks this up


def python_classifier_boolean_wrapper(audio_signal: np.ndarray, **kwargs):
    """python_classifier_boolean_wrapper (:2577-2598)."""
    rain_drop_count, _, _ = rain_detection_algo(audio_signal, **kwargs)
    if rain_drop_count > 0:
        return True
    if rain_drop_count == 0:
        return False
    return np.nan
