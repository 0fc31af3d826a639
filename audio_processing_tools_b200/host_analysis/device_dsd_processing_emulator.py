"""GPU twin of the drop-size-distribution emulator (reference:
host_analysis/device_dsd_processing_emulator.py, class ``DsdProcessingEmualtor`` :16-314).

Same constructor arguments, attributes (band indices, thresholds) and
``process_audio_data(audio_data, ts) -> list of 100-vectors`` (one per processed minute: 32 drop-size bins,
30 peak-frequency slots, 38 FFT energies).  ``process_audio_batch`` runs many clips in one GPU pass.
The spectra and the per-minute state machine run in CUDA behind ``apt_dsd_run_i16``; there is no CPU path.

Input scale: the reference feeds ``parse.pcm_to_float(sig)`` = int16 / 32768 in float64 (parse.py:670); this
twin takes the int16 samples themselves (or a float array that is exactly int16 / 32768) so that the device
works on the wire format.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Sequence

import numpy as np
from scipy.signal import get_window

from .. import _lib
from .._staging import upload_clips
from ..engine import AptError, _torch


def _as_int16(audio_data) -> np.ndarray:
    a = np.asarray(audio_data)
    if a.dtype == np.int16:
        return np.ascontiguousarray(a.reshape(-1))
    if np.issubdtype(a.dtype, np.floating):
        q = np.rint(a.astype(np.float64) * 32768.0)
        if not np.array_equal(q / 32768.0, a.astype(np.float64)) or q.min(initial=0) < -32768 or q.max(initial=0) > 32767:
            raise ValueError("float input must be int16 PCM scaled by 1/32768 (parse.pcm_to_float); "
                             "pass the int16 samples instead")
        return np.ascontiguousarray(q.astype(np.int16).reshape(-1))
    raise TypeError(f"unsupported audio dtype {a.dtype}")


class DsdProcessingEmualtor:
    def __init__(self, fs=11162, frame_length=512, hop_length=512, bwindow=False, ts=0, verbose=False, device=0):
        self.fs = fs
        self.frame_length = frame_length
        self.fft_n_bins = int(frame_length / 2)
        self.hop_length = hop_length
        self.apply_window = bwindow
        self.verbose = verbose
        self.dF = self.fs / self.frame_length
        self.loudness_bins, self.pft_bins, self.fft_bins = 32, 30, 38
        self.rain_chk_period_seconds, self.rain_chk_duration_seconds = 60, 3
        self.rain_energy_threshold = 0.6
        self.rain_low_freq, self.rain_high_freq = 400, 700
        self.rain_low_idx = int(self.rain_low_freq // self.dF) + 1
        self.rain_high_idx = int(self.rain_high_freq // self.dF)
        self.rain_log_base, self.rain_log_factor = 1.13, 0.6
        self.pft_low_freq, self.pft_high_freq = 100, 1500
        self.pft_low_idx = int(self.pft_low_freq // self.dF) + 1
        self.pft_high_idx = int(self.pft_high_freq // self.dF) - 1
        self.lwin_start, self.hwin_start = 300, 1000
        self.lwin_start_idx = int(self.lwin_start // self.dF)
        self.lwin_end_idx = self.lwin_start_idx + int(self.fft_bins // 2) - 1
        self.hwin_start_idx = int(self.hwin_start // self.dF)
        self.hwin_end_idx = self.hwin_start_idx + int(self.fft_bins // 2) - 1
        self.raining = True
        self._device = int(device)
        self._ctx = None

    def _context(self):
        if self._ctx is None:
            _torch()
            L = _lib.load()
            ctx = C.c_void_p()
            if L.apt_init(self._device, C.byref(ctx)) != 0:
                raise AptError("apt_init failed: is a B200 visible?")
            self._ctx = ctx
        return _lib.load(), self._ctx

    def process_audio_batch(self, clips: Sequence[np.ndarray], ts: Sequence[float]) -> List[List[np.ndarray]]:
        """Every clip starts from the constructor state (raining = True), as one fresh reference object per clip."""
        torch = _torch()
        L, ctx = self._context()
        pcm = [_as_int16(c) for c in clips]
        if len(pcm) != len(ts) or not pcm:
            raise ValueError("clips and ts must have the same non-zero length")
        lens = np.array([p.size for p in pcm], dtype=np.int64)
        max_minutes = max(1, int(max(math.ceil(n / (self.fs * 60)) for n in lens)))
        dev = torch.device("cuda", self._device)
        d_pcm = upload_clips(torch, dev, pcm, lens)            # pinned staging buffer, one asynchronous copy
        d_out = torch.zeros((len(pcm), max_minutes, 100), dtype=torch.float64, device=dev)
        d_n = torch.zeros(len(pcm), dtype=torch.int32, device=dev)
        prm = _lib.AptDsdParams()
        prm.fs, prm.frame_length, prm.hop_length = int(self.fs), int(self.frame_length), int(self.hop_length)
        prm.apply_window = int(bool(self.apply_window))
        win = np.ascontiguousarray(get_window("hann", int(self.frame_length)).astype(np.float64))
        prm.window = win.ctypes.data_as(C.c_void_p)
        tsa = np.ascontiguousarray(np.asarray(ts, dtype=np.float64))
        with _lib.device_timer(torch, "dsd", self._device):
            rc = L.apt_dsd_run_i16(ctx, C.byref(prm), len(pcm), lens.ctypes.data_as(C.POINTER(C.c_int64)),
                                   tsa.ctypes.data_as(C.POINTER(C.c_double)), d_pcm.data_ptr(), d_out.data_ptr(),
                                   d_n.data_ptr(), max_minutes, torch.cuda.current_stream(self._device).cuda_stream)
        if rc != 0:
            raise AptError(f"apt_dsd_run_i16 failed ({rc}): {L.apt_last_error(ctx).decode()}")
        out, n = d_out.cpu().numpy(), d_n.cpu().numpy()
        return [[out[c, m].copy() for m in range(int(n[c]))] for c in range(len(pcm))]

    def process_audio_data(self, audio_data, ts):
        if not self.raining:
            raise NotImplementedError("carrying the not-raining state into a new call is not implemented; "
                                      "create a new emulator per clip (as transform.process_audio_file_dsd does)")
        return self.process_audio_batch([audio_data], [ts])[0]
