"""Batch engine: drives libapt_b200.so for a list of clips.

PyTorch is used only as the carrier of device buffers and streams; all arithmetic happens in the
CUDA kernels behind the C ABI.  One ``BatchEngine`` per (config, GPU); plans are cached per tuple
of clip lengths so repeated batches of the same shape reuse scratch.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Any, Dict, Iterable, List, Optional, Sequence

import numpy as np

from . import _lib
from .config import NoiseProcessorConfig, ResolvedParams

# optional output planes -> (torch dtype name, shape builder(nF, nS, rp))
_PLANES = {
    "S": ("float32", lambda nF, nS, rp: (nF, rp.F, 2)),
    "P": ("float32", lambda nF, nS, rp: (nF, rp.F)),
    "det_noise_psd": ("float32", lambda nF, nS, rp: (nF, rp.K)),
    "det_noise_lag": ("float32", lambda nF, nS, rp: (nF, rp.K)),
    "D": ("float32", lambda nF, nS, rp: (nF, rp.K)),
    "noise_psd": ("float32", lambda nF, nS, rp: (nF, rp.K)),
    "mode_flux": ("float32", lambda nF, nS, rp: (rp.M, nF)),
    "norm_flux": ("float32", lambda nF, nS, rp: (rp.M, nF)),
    "score": ("float32", lambda nF, nS, rp: (nF,)),
    "td": ("float32", lambda nF, nS, rp: (_lib.N_TD, nF)),
    "raw": ("float32", lambda nF, nS, rp: (_lib.N_RAW, nF)),
    "band_energy": ("float32", lambda nF, nS, rp: (rp.M + 1, nF)),
    "gate": ("uint8", lambda nF, nS, rp: (nF,)),
    "x_td": ("float32", lambda nF, nS, rp: (nS,)),
    "G": ("float32", lambda nF, nS, rp: (nF, rp.K)),
    "ratio_med": ("float32", lambda nF, nS, rp: (nF,)),
    "S_hat": ("float32", lambda nF, nS, rp: (nF, rp.F, 2)),
    "y": ("float32", lambda nF, nS, rp: (nS,)),
    "peak_ratio": ("float32", lambda nF, nS, rp: (nF,)),
    "peak_gate_score": ("float32", lambda nF, nS, rp: (nF,)),
    "peak_valid_count": ("int32", lambda nF, nS, rp: (nF,)),
    "peak_count_by_mode": ("int32", lambda nF, nS, rp: (rp.M, nF)),
    "td_fast_crest": ("float32", lambda nF, nS, rp: (nF,)),
    "snr_mode": ("float32", lambda nF, nS, rp: (nF,)),
    "snr_gate": ("float32", lambda nF, nS, rp: (nF,)),
}
_CORE = {
    "frame_class": ("int8", lambda nF, nC: (nF,)),
    "rain_conf": ("float32", lambda nF, nC: (nF,)),
    "noise_conf": ("float32", lambda nF, nC: (nF,)),
    "event_idx": ("int32", lambda nF, nC: (nF,)),
    "event_count": ("int32", lambda nF, nC: (nC,)),
    "clip_stats": ("float32", lambda nF, nC: (nC, _lib.N_STATS)),
}
STAT_NAMES = ("clip_id", "rain_frame_count", "clip_rain_fraction", "clip_is_rain", "clip_rain_conf",
              "median_rain_conf", "mean_noise_floor_db", "median_noise_floor_db")


class AptError(RuntimeError):
    pass


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise AptError("no CUDA device visible: the audio_processing_tools_b200 hot path runs only on the GPU "
                       "(there is no CPU fallback)")
    return torch


class Plan:
    def __init__(self, engine: "BatchEngine", lengths: Sequence[int]):
        L = engine.L
        self.engine = engine
        self.n_clips = len(lengths)
        lens = (C.c_int64 * self.n_clips)(*[int(n) for n in lengths])
        h = C.c_void_p()
        rc = L.apt_plan_create(engine.ctx, C.byref(engine.rp.c), self.n_clips, lens, C.byref(h))
        if rc != 0:
            raise AptError(f"apt_plan_create failed ({rc}): {L.apt_last_error(engine.ctx).decode()}")
        self.h = h
        so = (C.c_int64 * (self.n_clips + 1))()
        fo = (C.c_int64 * (self.n_clips + 1))()
        L.apt_plan_offsets(h, so, fo)
        self.sample_off = np.frombuffer(so, dtype=np.int64).copy()
        self.frame_off = np.frombuffer(fo, dtype=np.int64).copy()
        self.nS = int(self.sample_off[-1])
        self.nF = int(self.frame_off[-1])

    def close(self):
        if self.h:
            self.engine.L.apt_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BatchEngine:
    """GPU engine for one resolved configuration."""

    def __init__(self, cfg: NoiseProcessorConfig, sample_rate: int, *, device: int = 0,
                 clip_rain_min_frames: int = 1, fft_f64: bool = True, max_plans: int = 4):
        self.L = _lib.load()
        torch = _torch()
        self.torch = torch
        self.device = int(device)
        self.rp = ResolvedParams(cfg, sample_rate, clip_rain_min_frames=clip_rain_min_frames, fft_f64=fft_f64)
        ctx = C.c_void_p()
        rc = self.L.apt_init(self.device, C.byref(ctx))
        if rc != 0:
            raise AptError(f"apt_init({device}) failed with {rc}: is a B200 visible?")
        self.ctx = ctx
        self._plans: Dict[tuple, Plan] = {}
        self._max_plans = max_plans
        self.last_launches = 0
        self.last_host_call_s = 0.0      # seconds inside the last apt_run_host_clips call

    # ------------------------------------------------------------------
    def plan_for(self, lengths: Sequence[int]) -> Plan:
        key = tuple(int(n) for n in lengths)
        pl = self._plans.get(key)
        if pl is None:
            if len(self._plans) >= self._max_plans:
                _, old = self._plans.popitem()
                old.close()
            pl = Plan(self, key)
            self._plans[key] = pl
        return pl

    def close(self):
        for pl in self._plans.values():
            pl.close()
        self._plans.clear()
        if getattr(self, "ctx", None):
            self.L.apt_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------
    def alloc_outputs(self, plan: Plan, want: Iterable[str] = (), full: bool = True):
        torch = self.torch
        dev = torch.device("cuda", self.device)
        bufs: Dict[str, Any] = {}
        if full:
            for name, (dt, shp) in _CORE.items():
                bufs[name] = torch.empty(shp(plan.nF, plan.n_clips), dtype=getattr(torch, dt), device=dev)
        for name in want:
            dt, shp = _PLANES[name]
            bufs[name] = torch.empty(shp(plan.nF, plan.nS, self.rp), dtype=getattr(torch, dt), device=dev)
        return bufs

    def _out_struct(self, bufs):
        o = _lib.AptOut()
        for name in _lib.OUT_FIELDS:
            t = bufs.get(name)
            setattr(o, name, t.data_ptr() if t is not None else None)
        return o

    def run_device(self, plan: Plan, pcm_dev, bufs, *, full: bool = True, stream=None):
        """Enqueue one pass over a device-resident concatenated PCM tensor (int16 or float32)."""
        torch = self.torch
        if pcm_dev.numel() != plan.nS:
            raise AptError(f"PCM tensor has {pcm_dev.numel()} samples, plan expects {plan.nS}")
        o = self._out_struct(bufs)
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        stages = _lib.STAGE_FULL if full else _lib.STAGE_FEATURES
        if pcm_dev.dtype == torch.int16:
            rc = self.L.apt_run_i16(plan.h, stages, pcm_dev.data_ptr(), C.byref(o), st.cuda_stream)
        elif pcm_dev.dtype == torch.float32:
            rc = self.L.apt_run_f32(plan.h, stages, pcm_dev.data_ptr(), C.byref(o), st.cuda_stream)
        else:
            raise TypeError(f"PCM dtype {pcm_dev.dtype} unsupported (int16 or float32)")
        if rc != 0:
            raise AptError(f"apt_run failed ({rc}): {self.L.apt_last_error(self.ctx).decode()}")
        self.last_launches = int(self.L.apt_plan_last_launches(plan.h))
        return bufs

    def run_host_i16(self, plan: Plan, pcm_host: np.ndarray, outs: Dict[str, np.ndarray]):
        """End-to-end path with host buffers (pinned recommended): H2D, full pipeline, D2H."""
        assert pcm_host.dtype == np.int16 and pcm_host.size == plan.nS
        ptr = lambda k: outs[k].ctypes.data if outs.get(k) is not None else None
        rc = self.L.apt_run_host_i16(plan.h, pcm_host.ctypes.data, ptr("frame_class"), ptr("rain_conf"),
                                     ptr("noise_conf"), ptr("event_idx"), ptr("event_count"), ptr("clip_stats"))
        if rc != 0:
            raise AptError(f"apt_run_host_i16 failed ({rc}): {self.L.apt_last_error(self.ctx).decode()}")
        self.last_launches = int(self.L.apt_plan_last_launches(plan.h))
        return outs

    def run_host_clips(self, clips: List[np.ndarray], *, event_idx: bool = False):
        """The plugin's default-flags path: host clips (all int16 or all float32, pageable memory is fine) ->
        apt_run_host_clips (pinned staging ring, H2D / compute / D2H pipelined over clip groups) -> results in
        freshly allocated pinned host arrays that the caller owns.  Returns (plan, outs)."""
        torch = self.torch
        first = clips[0]
        if first.dtype not in (np.int16, np.float32):
            raise TypeError(f"clip dtype {first.dtype} unsupported on the host path (int16 or float32)")
        arrs = []
        for c in clips:
            if c.dtype != first.dtype or c.ndim != 1:
                raise TypeError("all clips of a batch must be 1-D arrays of one dtype")
            arrs.append(c if c.flags.c_contiguous else np.ascontiguousarray(c))
        plan = self.plan_for([a.size for a in arrs])
        n, nF = plan.n_clips, plan.nF
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])

        def pin(shape, dt):
            return torch.empty(shape, dtype=dt, pin_memory=True).numpy()

        outs = {"frame_class": pin((nF,), torch.int8), "rain_conf": pin((nF,), torch.float32),
                "noise_conf": pin((nF,), torch.float32), "event_count": pin((n,), torch.int32),
                "clip_stats": pin((n, _lib.N_STATS), torch.float32),
                "event_idx": pin((nF,), torch.int32) if event_idx else None}
        ptr = lambda k: outs[k].ctypes.data if outs.get(k) is not None else None
        t0 = time.perf_counter()
        rc = self.L.apt_run_host_clips(plan.h, ptrs, int(first.dtype == np.float32), ptr("frame_class"), ptr("rain_conf"),
                                       ptr("noise_conf"), ptr("event_idx"), ptr("event_count"), ptr("clip_stats"))
        self.last_host_call_s = time.perf_counter() - t0
        if rc != 0:
            raise AptError(f"apt_run_host_clips failed ({rc}): {self.L.apt_last_error(self.ctx).decode()}")
        self.last_launches = int(self.L.apt_plan_last_launches(plan.h))
        return plan, outs

    # ------------------------------------------------------------------
    def run_clips(self, clips: List[np.ndarray], want: Iterable[str] = (), *, full: bool = True):
        """Convenience: host clips (all int16 or all float) -> per-batch numpy results."""
        torch = self.torch
        if not clips:
            return None
        first = np.asarray(clips[0])
        if first.dtype == np.int16:
            cat = np.concatenate([np.ascontiguousarray(c, dtype=np.int16).reshape(-1) for c in clips])
        else:
            cat = np.concatenate([np.ascontiguousarray(c, dtype=np.float32).reshape(-1) for c in clips])
        plan = self.plan_for([np.asarray(c).size for c in clips])
        dev = torch.device("cuda", self.device)
        pcm = torch.from_numpy(cat).to(dev, non_blocking=False)
        bufs = self.alloc_outputs(plan, want, full=full)
        self.run_device(plan, pcm, bufs, full=full)
        torch.cuda.synchronize(self.device)
        return plan, {k: v.cpu().numpy() for k, v in bufs.items()}
