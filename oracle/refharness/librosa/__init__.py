"""Minimal stand-in for librosa==0.11.0 (pinned in the reference's uv.lock:813-814).

TEST INFRASTRUCTURE ONLY.  It exists so that the UNMODIFIED reference under
/root/reference can be imported in the build container (librosa is not
installable offline) to generate golden vectors (oracle/make_golden.py).
It restates only the public librosa semantics that the reference calls:

* stft          -- edge/rain_signal_processor.py:818-825, edge/dsp_rain_detection.py:2283
* istft         -- edge/rain_signal_processor.py:1115-1122 (compute_output_audio only)
* fft_frequencies / frames_to_time -- edge/rain_signal_processor.py:827-828
* amplitude_to_db / power_to_db    -- edge/dsp_rain_detection.py:2337-2338
* load / resample                  -- audio_io.py:109,408 (host I/O; raise here)

No reference test pins results at this boundary => "parity unpinned" here
(see DESIGN.md).  Semantics restated from librosa 0.11 public behaviour:
center=True pads n_fft//2 zeros each side (pad_mode="constant"); window is
scipy.signal.get_window("hann", n_fft, fftbins=True) in float64; frame count
1 + len(y)//hop; window*frames promotes to float64; scipy.fft.rfft along the
frame axis; result stored in a complex64 Fortran-ordered matrix for float32
input.
"""
from __future__ import annotations

import numpy as np
import scipy.fft
import scipy.signal

__version__ = "0.11.0-standin"


def _window(window, win_length, n_fft):
    w = scipy.signal.get_window(window, win_length, fftbins=True)
    if win_length < n_fft:
        lpad = (n_fft - win_length) // 2
        w = np.pad(w, (lpad, n_fft - win_length - lpad))
    return w


def stft(y, *, n_fft=2048, hop_length=None, win_length=None, window="hann",
         center=True, dtype=None, pad_mode="constant", out=None):
    y = np.asarray(y)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = win_length // 4
    fft_window = _window(window, win_length, n_fft).reshape(-1, 1)
    if center:
        if pad_mode != "constant":
            raise NotImplementedError("stand-in supports pad_mode='constant' only")
        y = np.pad(y, (n_fft // 2, n_fft // 2), mode="constant")
    if y.shape[-1] < n_fft:
        raise ValueError("input too short")
    n_frames = 1 + (y.shape[-1] - n_fft) // hop_length
    frames = np.lib.stride_tricks.as_strided(
        y, shape=(n_fft, n_frames),
        strides=(y.strides[-1], hop_length * y.strides[-1]), writeable=False)
    if dtype is None:
        dtype = np.complex64 if y.dtype == np.float32 else np.complex128
    S = np.empty((1 + n_fft // 2, n_frames), dtype=dtype, order="F")
    # librosa processes column blocks bounded by MAX_MEM_BLOCK; blocks do not
    # change the arithmetic of a per-column FFT.
    blk = max(1, (2 ** 18) // max(1, n_fft))
    for b0 in range(0, n_frames, blk):
        b1 = min(n_frames, b0 + blk)
        S[:, b0:b1] = scipy.fft.rfft(fft_window * frames[:, b0:b1], axis=0)
    return S


def istft(stft_matrix, *, hop_length=None, win_length=None, n_fft=None,
          window="hann", center=True, dtype=None, length=None, out=None):
    S = np.asarray(stft_matrix)
    if n_fft is None:
        n_fft = 2 * (S.shape[0] - 1)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = win_length // 4
    w = _window(window, win_length, n_fft)
    n_frames = S.shape[1]
    if dtype is None:
        dtype = np.float32 if S.dtype == np.complex64 else np.float64
    expected = n_fft + hop_length * (n_frames - 1)
    y = np.zeros(expected, dtype=np.float64)
    wss = np.zeros(expected, dtype=np.float64)
    ytmp = w.reshape(-1, 1) * scipy.fft.irfft(S, n=n_fft, axis=0)
    wsq = w ** 2
    for t in range(n_frames):
        s = t * hop_length
        y[s:s + n_fft] += ytmp[:, t]
        wss[s:s + n_fft] += wsq
    tiny = np.finfo(np.float32).tiny
    nz = wss > tiny
    y[nz] /= wss[nz]
    if center:
        y = y[n_fft // 2:]
    if length is not None:
        if y.shape[0] >= length:
            y = y[:length]
        else:
            y = np.pad(y, (0, length - y.shape[0]))
    elif center:
        y = y[: y.shape[0] - n_fft // 2]
    return y.astype(dtype)


def fft_frequencies(*, sr=22050, n_fft=2048):
    return np.fft.rfftfreq(n=n_fft, d=1.0 / sr)


def frames_to_samples(frames, *, hop_length=512, n_fft=None):
    offset = int(n_fft // 2) if n_fft is not None else 0
    return (np.asanyarray(frames) * hop_length + offset).astype(int)


def frames_to_time(frames, *, sr=22050, hop_length=512, n_fft=None):
    return frames_to_samples(frames, hop_length=hop_length, n_fft=n_fft) / float(sr)


def power_to_db(S, *, ref=1.0, amin=1e-10, top_db=80.0):
    S = np.asarray(S)
    magnitude = np.abs(S) if np.iscomplexobj(S) else S
    ref_value = ref(magnitude) if callable(ref) else np.abs(ref)
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def amplitude_to_db(S, *, ref=1.0, amin=1e-5, top_db=80.0):
    S = np.asarray(S)
    magnitude = np.abs(S)
    ref_value = ref(magnitude) if callable(ref) else np.abs(ref)
    power = np.square(magnitude, out=magnitude)
    return power_to_db(power, ref=ref_value ** 2, amin=amin ** 2, top_db=top_db)


def load(*a, **k):
    raise NotImplementedError("librosa stand-in: load() is host I/O, not available")


def resample(*a, **k):
    raise NotImplementedError("librosa stand-in: resample() is host I/O, not available")
