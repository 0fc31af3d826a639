"""GPU twin of ``BandNoiseEstimatorProcessor`` (reference: edge/band_noise_processor.py:14-281).

Same ``name`` / ``mode`` / ``run(audio_data, params) -> (results, state)`` contract, the same configuration keys
(estimator attributes, ``det.*`` dotted detector overrides, ``sample_rate`` / ``fs``), the same result and state
keys; ``run_batch`` processes a list of clips in one GPU pass.  The per-frame estimator loop runs in CUDA behind
``apt_bne_run`` (float64 configuration); there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List, Sequence, Tuple

import numpy as np
import scipy.signal as spsig

from .. import _lib
from .._staging import upload_clips
from ..engine import AptError, _torch
from .band_noise_estimator import BandNoiseEstimatorConfig, NoiseFrameDetectorConfig, db_to_ratio, hz_to_bin

_FRAME_KEYS = ("M_band", "E_band", "N_E", "N_E_raw", "G_mag", "M_clean", "noise_effective_q", "M_band_fft", "E_band_fft", "E_hpf")
_STAT_KEYS = ("noise_energy_sum", "rain_energy_sum", "total_energy_sum", "noise_frame_count", "rain_frame_count",
              "total_frame_count", "noise_buffer_valid_count", "noise_buffer_min_valid_count",
              "noise_buffer_underflow_frame_count", "frames_since_noise_update", "noise_learned_subframe_count",
              "noise_replenish_count", "noise_effective_q")


def _warmup_samples(sos_list, frame_len: int) -> int:
    """Samples after which the cascade's memory of its initial state is below 1e-17 (largest pole radius)."""
    r = 0.0
    for sos in sos_list:
        if sos is None:
            continue
        for sec in sos:
            r = max(r, float(np.max(np.abs(np.roots([1.0, sec[4], sec[5]])))))
    if r <= 0.0:
        return 0
    if r >= 0.99999:
        raise NotImplementedError("filter poles too close to the unit circle for segmented streaming")
    n = int(np.ceil(np.log(1e-17) / np.log(r))) * 2     # cascaded sections: polynomial-times-exponential tails
    return int(np.ceil(n / frame_len) * frame_len)


class BandNoiseEstimatorProcessor:
    def __init__(self, name: str = "band_noise", mode: str = "fft", device: int = 0):
        self.name = name
        self.mode = (mode or "fft").lower().strip()
        self._device = int(device)
        self._ctx = None

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_ctx"] = None
        return d

    # -- configuration (edge/band_noise_processor.py:32-76)
    def _build_config(self, params: Dict[str, Any]) -> BandNoiseEstimatorConfig:
        cfg = BandNoiseEstimatorConfig()
        for k, v in params.items():
            if k.startswith("det."):
                subk = k.split(".", 1)[1]
                if hasattr(cfg.det, subk):
                    setattr(cfg.det, subk, v)
                continue
            if hasattr(cfg, k):
                if k == "dtype":
                    v = {"float32": np.float32, "np.float32": np.float32, "float64": np.float64, "np.float64": np.float64}.get(v, v)
                setattr(cfg, k, v)
        if "sample_rate" in params:
            cfg.fs = int(params["sample_rate"])
        elif "fs" in params:
            cfg.fs = int(params["fs"])
        cfg.det.fs = int(cfg.fs)
        cfg.det.n_fft = int(cfg.frame_len)
        cfg.validate()
        return cfg

    def _resolve(self, cfg: BandNoiseEstimatorConfig) -> "_lib.AptBneParams":
        if cfg.dtype is not np.float64:
            raise NotImplementedError("dtype=float32 is not implemented on the CUDA path (float64 is the reference default)")
        if int(cfg.subhop) != int(cfg.subframe_len):
            raise NotImplementedError("overlapping subframes (subhop != subframe_len) are not implemented on the CUDA path")
        P = _lib.AptBneParams()
        N = int(cfg.frame_len)
        P.fs, P.N, P.sub_len = int(cfg.fs), N, int(cfg.subframe_len)
        P.S = 1 + (N - int(cfg.subframe_len)) // int(cfg.subhop)
        nyq = 0.5 * cfg.fs
        hpf = None
        if cfg.hp_cutoff_hz > 0:
            hpf = spsig.butter(cfg.hp_order, np.clip(cfg.hp_cutoff_hz / nyq, 1e-6, 0.999), btype="highpass", output="sos")
        lo, hi = cfg.band_hz
        w1, w2 = np.clip(lo / nyq, 1e-6, 0.999), np.clip(hi / nyq, 1e-6, 0.999)
        if w2 <= w1:
            w2 = min(0.999, w1 + 1e-3)
        bpf = spsig.butter(cfg.bpf_order, [w1, w2], btype="bandpass", output="sos")
        for sos, dst_s, dst_z, cnt in ((hpf, P.sos_h, P.zi_h, "ns_h"), (bpf, P.sos_b, P.zi_b, "ns_b")):
            n = 0 if sos is None else int(sos.shape[0])
            if n > _lib.BNE_MAX_SOS:
                raise NotImplementedError(f"filter with {n} second-order sections")
            setattr(P, cnt, n)
            if sos is not None:
                zi = spsig.sosfilt_zi(sos)
                for s in range(n):
                    for j in range(6):
                        dst_s[s][j] = float(sos[s, j])
                    dst_z[s][0], dst_z[s][1] = float(zi[s, 0]), float(zi[s, 1])
        P.warm = _warmup_samples((hpf, bpf), N)
        det = cfg.det
        bands = tuple(det.rain_bands_hz)
        if len(bands) > _lib.BNE_MAX_BANDS:
            raise NotImplementedError(f"more than {_lib.BNE_MAX_BANDS} rain bands")
        P.n_bands = len(bands)
        for i, (f0, f1) in enumerate(bands):
            P.band_b0[i], P.band_b1[i] = hz_to_bin(f0, det.fs, det.n_fft), hz_to_bin(f1, det.fs, det.n_fft)
        P.prim_b0, P.prim_b1 = hz_to_bin(det.primary_hz[0], det.fs, det.n_fft), hz_to_bin(det.primary_hz[1], det.fs, det.n_fft)
        freqs = np.fft.rfftfreq(N, d=1.0 / cfg.fs)
        idx = np.flatnonzero((freqs >= lo) & (freqs <= hi))
        P.mask_b0, P.mask_b1 = (int(idx[0]), int(idx[-1])) if idx.size else (1, 0)
        P.M_ratio, P.N_ratio, P.D_ratio = db_to_ratio(det.M_db), db_to_ratio(det.N_db), db_to_ratio(det.D_db)
        P.band_rise_db, P.excess_rise_db = float(det.band_rise_db), float(det.excess_rise_db)
        P.min_Ehpf, P.min_Eband, P.dE_thr = float(det.min_Ehpf), float(det.min_Eband), float(det.dE_over_Ehpf_thr)
        P.k_subframes, P.use_dE, P.use_D = int(det.k_subframes), int(bool(det.use_dE_over_Ehpf)), int(bool(det.use_D_trigger))
        P.W, P.W_min, P.ttl = int(cfg.W), int(cfg.W_min), int(cfg.noise_buffer_ttl_frames)
        P.smooth = int(bool(cfg.smooth_N_E))
        P.learn_all = int(bool(cfg.force_learn_all) or bool(cfg.learn_during_rain))
        P.replenish = int(bool(cfg.noise_replenish_from_all_subframes))
        P.replenish_only_not_full = int(bool(cfg.noise_replenish_only_when_buffer_not_full))
        P.q_adapt = int(bool(cfg.noise_q_adapt_enable))
        P.q, P.ema_alpha, P.beta, P.gain_floor, P.eps = float(cfg.q), float(cfg.ema_alpha), float(cfg.beta), float(cfg.gain_floor), float(cfg.eps)
        P.att_dry, P.att_wet, P.release = float(cfg.ne_attack_alpha_dry), float(cfg.ne_attack_alpha_wet), float(cfg.ne_release_alpha)
        P.repl_q, P.q_repl_alpha, P.q_norm_alpha = float(cfg.noise_replenish_q), float(cfg.noise_q_replenish_alpha), float(cfg.noise_q_normal_alpha)
        return P

    def _context(self):
        if self._ctx is None:
            _torch()
            L = _lib.load()
            ctx = C.c_void_p()
            if L.apt_init(self._device, C.byref(ctx)) != 0:
                raise AptError("apt_init failed: is a B200 visible?")
            self._ctx = ctx
        return _lib.load(), self._ctx

    # -- framework entry points
    def run(self, audio_data: np.ndarray, params: Dict[str, Any]) -> Tuple[Dict[str, Any], Dict[str, Any]]:
        return self.run_batch([audio_data], params)[0]

    def run_batch(self, audio_list: Sequence[np.ndarray], params: Dict[str, Any]) -> List[Tuple[Dict[str, Any], Dict[str, Any]]]:
        cfg = self._build_config(params)
        N, fs = int(cfg.frame_len), int(cfg.fs)
        hop = int(params.get("hop", N))
        if hop <= 0:
            raise ValueError("hop must be positive")
        if hop != N:
            raise ValueError("BandNoiseEstimatorProcessor requires hop == frame_len because BandNoiseEstimator keeps "
                             f"streaming IIR filter state across frames. Got hop={hop}, frame_len={N}.")
        P = self._resolve(cfg)
        S = int(P.S)
        clips = []
        for a in audio_list:
            x = np.asarray(a)
            if x.ndim != 1 or x.size == 0:
                raise ValueError("audio_data must be non-empty mono ndarray")
            clips.append(x if x.dtype == np.int16 else np.asarray(x, dtype=np.float32))
        is_f32 = clips[0].dtype != np.int16
        if any((c.dtype != np.int16) != is_f32 for c in clips):
            raise TypeError("a batch must be all int16 or all float")
        torch = _torch()
        L, ctx = self._context()
        lens = np.array([c.size for c in clips], dtype=np.int64)
        nfr = lens // N
        nF = int(nfr.sum())
        dev = torch.device("cuda", self._device)
        d_pcm = upload_clips(torch, dev, clips, lens)
        d_fo = torch.zeros((max(nF, 1), _lib.BNE_FRAME_F), dtype=torch.float64, device=dev)
        d_mask = torch.zeros(max(nF, 1), dtype=torch.uint8, device=dev)
        d_sub = torch.zeros((max(nF, 1), _lib.BNE_MAX_S), dtype=torch.float64, device=dev)
        d_st = torch.zeros((len(clips), _lib.BNE_STATS), dtype=torch.float64, device=dev)
        with _lib.device_timer(torch, "bne", self._device):
            rc = L.apt_bne_run(ctx, C.byref(P), len(clips), lens.ctypes.data_as(C.POINTER(C.c_int64)), d_pcm.data_ptr(), int(is_f32),
                               d_fo.data_ptr(), d_mask.data_ptr(), d_sub.data_ptr(), d_st.data_ptr(),
                               torch.cuda.current_stream(self._device).cuda_stream)
        if rc != 0:
            raise AptError(f"apt_bne_run failed ({rc}): {L.apt_last_error(ctx).decode()}")
        # per-frame columns leave the device column-major, so that a clip's column is one contiguous run of the batch array
        fo_t = d_fo.t().contiguous().cpu().numpy()
        mask, st = d_mask.cpu().numpy(), d_st.cpu().numpy()
        sub = d_sub[:, :S].contiguous().cpu().numpy()
        return self._package_batch(cfg, clips, params, fo_t, mask, sub, st, nfr, S, N, fs)

    @staticmethod
    def _median(a: np.ndarray) -> float:
        """np.median of a 1-D float array without its per-call overhead (same two-element mean for even sizes)."""
        n = a.size
        h = n >> 1
        if n & 1:
            return float(np.partition(a, h)[h])
        part = np.partition(a, (h - 1, h))
        r = float((part[h - 1] + part[h]) / 2.0)
        return r if r == r or not np.isnan(a).any() else float("nan")

    def _package_batch(self, cfg, clips, params, fo_t, mask, sub, st, nfr, S, N, fs):
        """Result / state dictionaries of every clip (edge/band_noise_processor.py:200-281), built from batch-wide arrays."""
        dtype = cfg.dtype
        nF = int(nfr.sum())
        submask_all = ((mask[:nF, None] >> np.arange(S, dtype=np.uint8)[None, :]) & 1).astype(bool)
        fft_rain_all = fo_t[11, :nF] > 0.5
        nsub_all = np.repeat(fo_t[10, :nF, None], S, axis=1).astype(dtype, copy=False)
        sub_all = sub[:nF].astype(dtype, copy=False)
        want_audio = bool(params.get("include_audio_in_state", False))
        outs, f0 = [], 0
        nan = np.nan
        for c, x in enumerate(clips):
            n = int(nfr[c])
            f1 = f0 + n
            if n == 0:
                energy = {k: (0.0 if "sum" in k or k == "noise_effective_q" else 0) for k in _STAT_KEYS
                          if k not in ("noise_buffer_min_valid_count", "noise_buffer_underflow_frame_count", "frames_since_noise_update")}
                energy["noise_effective_q"] = float(cfg.q)
                energy.update(noise_energy_mean=0.0, rain_energy_mean=0.0, total_energy_mean=0.0)
            else:
                row = st[c]
                energy = {k: (float(row[i]) if ("sum" in k or k == "noise_effective_q") else int(row[i])) for i, k in enumerate(_STAT_KEYS)}
                energy["noise_energy_mean"] = energy["noise_energy_sum"] / max(1, energy["noise_frame_count"])
                energy["rain_energy_mean"] = energy["rain_energy_sum"] / max(1, energy["rain_frame_count"])
                energy["total_energy_mean"] = energy["total_energy_sum"] / max(1, energy["total_frame_count"])
            blk = fo_t[:len(_FRAME_KEYS), f0:f1].astype(dtype, copy=True)     # one block per clip, a row per series
            cols = {k: blk[i] for i, k in enumerate(_FRAME_KEYS)}
            fft_rain = fft_rain_all[f0:f1].copy()
            results = {
                "processor": self.name, "mode": self.mode, "n_frames": n,
                "M_clean_med": self._median(cols["M_clean"]) if n else nan,
                "noise_E_med": self._median(cols["N_E"]) if n else nan,
                "gain_med": self._median(cols["G_mag"]) if n else nan,
                "noise_effective_q_last": float(cols["noise_effective_q"][-1]) if n else nan,
                "noise_effective_q_med": self._median(cols["noise_effective_q"]) if n else nan,
                "fft_rain_frac": float(np.mean(fft_rain)) if n else nan,
                **{f"energy_stats__{k}": v for k, v in energy.items()},
            }
            state: Dict[str, Any] = {
                "processor": self.name, "mode": self.mode, "times_s": (np.arange(n, dtype=np.float64) * N) / fs,
                **cols,
                "subE": sub_all[f0:f1].copy(),
                "N_sub": nsub_all[f0:f1].copy(),
                "rain_submask": submask_all[f0:f1].copy(),
                "fft_rain_frame": fft_rain,
                "config": cfg, "energy_stats": energy,
            }
            if want_audio:
                state["x_in"] = np.asarray(x, dtype=dtype).copy() if x.dtype != np.int16 else (x.astype(np.float32) / np.float32(32767.0)).astype(dtype)
            outs.append((results, state))
            f0 = f1
        return outs
