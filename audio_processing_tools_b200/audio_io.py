"""PCM conversion helpers (reference: audio_io.py:34-120).  Only the converters that feed the hot
path are provided; key discovery / S3 / DB loading are host I/O outside this package's scope."""
from __future__ import annotations

from typing import Optional

import numpy as np


def safe_to_float(data, bytes_per_sample: int = 2, signed: bool = True) -> np.ndarray:
    """int16 PCM (bytes or array) -> float32 / 32767; float arrays are clipped to [-1, 1].
    The CUDA path performs the same conversion on the device when handed int16 (apt_run_i16)."""
    if isinstance(data, (bytes, bytearray, memoryview)):
        if bytes_per_sample != 2 or not signed:
            raise ValueError("Only 16-bit signed PCM input is supported.")
        arr = np.frombuffer(data, dtype="<i2")
    else:
        arr = np.asarray(data)
    if np.issubdtype(arr.dtype, np.floating):
        return np.clip(arr.astype(np.float32, copy=False), -1.0, 1.0)
    if arr.dtype != np.int16:
        raise ValueError(f"Unsupported dtype {arr.dtype}; expected int16 or float.")
    return arr.astype(np.float32) / np.float32(32767.0)


def ensure_mono_len_sr(y: np.ndarray, sr_in: int, sr_out: int, duration_s: float) -> Optional[np.ndarray]:
    y = np.asarray(y)
    if y.ndim == 2:
        y = y.mean(axis=0) if y.shape[0] < y.shape[1] else y.mean(axis=1)
    if sr_in != sr_out:
        raise NotImplementedError("resampling is host I/O (librosa.resample in the reference); "
                                  "feed clips at the processing sample rate")
    need = int(sr_out * duration_s)
    if y.size < need:
        return None
    return np.clip(y[:need].astype(np.float32, copy=False), -1.0, 1.0)
