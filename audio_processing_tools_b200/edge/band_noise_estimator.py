"""Configuration of the band noise estimator (reference: edge/band_noise_estimator.py,
``NoiseFrameDetectorConfig`` :56-104, ``BandNoiseEstimatorConfig`` :414-513 incl. ``validate``).
The estimator itself runs on the GPU (csrc/apt_bne.cuh) behind ``apt_bne_run``; see
``band_noise_processor.BandNoiseEstimatorProcessor``."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Tuple

import numpy as np

EPS = 1e-12


def hz_to_bin(f_hz: float, fs: float, n_fft: int) -> int:
    return int(np.clip(np.round(f_hz * n_fft / fs), 0, n_fft // 2))


def db_to_ratio(db: float) -> float:
    return 10.0 ** (db / 10.0)


@dataclass
class NoiseFrameDetectorConfig:
    fs: int = 11162
    n_fft: int = 512
    M_db: float = 6.0
    N_db: float = 3.0
    primary_hz: Tuple[float, float] = (450.0, 650.0)
    rain_bands_hz: Tuple[Tuple[float, float], ...] = ((450.0, 650.0), (800.0, 1050.0), (1500.0, 1800.0),
                                                      (2350.0, 2550.0), (3150.0, 3350.0))
    k_subframes: int = 2
    band_rise_db: float = 6.0
    excess_rise_db: float = 3.0
    min_Ehpf: float = 1e-10
    min_Eband: float = 1e-12
    use_dE_over_Ehpf: bool = False
    dE_over_Ehpf_thr: float = 0.08
    use_D_trigger: bool = False
    D_db: float = 6.0


@dataclass
class BandNoiseEstimatorConfig:
    fs: int = 11162
    frame_len: int = 512
    dtype: type = np.float64
    hp_cutoff_hz: float = 350.0
    hp_order: int = 4
    band_hz: Tuple[float, float] = (400.0, 700.0)
    bpf_order: int = 4
    subframe_len: int = 128
    subhop: int = 128
    W: int = 30
    W_min: int = 10
    noise_buffer_ttl_frames: int = 200
    q: float = 0.3
    ema_alpha: float = 1
    beta: float = 1.0
    gain_floor: float = 0.10
    eps: float = 1e-12
    ne_attack_alpha_dry: float = 0.15
    ne_attack_alpha_wet: float = 0.02
    ne_release_alpha: float = 0.25
    smooth_N_E: bool = False
    learn_during_rain: bool = False
    force_learn_all: bool = False
    noise_replenish_from_all_subframes: bool = False
    noise_replenish_q: float = 0.20
    noise_replenish_only_when_buffer_not_full: bool = True
    noise_q_adapt_enable: bool = True
    noise_q_replenish_alpha: float = 0.2
    noise_q_normal_alpha: float = 0.1
    det: NoiseFrameDetectorConfig = field(default_factory=NoiseFrameDetectorConfig)

    def validate(self) -> None:
        if self.dtype not in (np.float32, np.float64):
            raise ValueError("dtype must be np.float32 or np.float64")
        if int(self.det.n_fft) != int(self.frame_len):
            raise ValueError("det.n_fft must match frame_len so FFT diagnostics and FFT rain detection use the same spectrum")
        if self.frame_len % self.subframe_len != 0:
            raise ValueError("subframe_len must divide frame_len")
        if not (0.0 < self.q < 1.0):
            raise ValueError("q must be in (0,1)")
        if not (0.0 < self.noise_replenish_q < 1.0):
            raise ValueError("noise_replenish_q must be in (0,1)")
        if not (0.0 < self.noise_q_replenish_alpha <= 1.0):
            raise ValueError("noise_q_replenish_alpha must be in (0,1]")
        if not (0.0 < self.noise_q_normal_alpha <= 1.0):
            raise ValueError("noise_q_normal_alpha must be in (0,1]")
        if self.W <= 0 or self.W_min < 0 or self.W_min > self.W:
            raise ValueError("Need W>0 and 0<=W_min<=W")
        if self.noise_buffer_ttl_frames < 0:
            raise ValueError("noise_buffer_ttl_frames must be >= 0")
        lo, hi = self.band_hz
        if not (0 < lo < hi < 0.5 * self.fs):
            raise ValueError("band_hz out of range")
        if not (0.0 < self.ema_alpha <= 1.0):
            raise ValueError("ema_alpha must be in (0, 1]")
        if not (isinstance(self.subhop, int) and self.subhop > 0):
            raise ValueError("subhop must be a positive integer")
        if self.frame_len < self.subframe_len:
            raise ValueError("frame_len must be >= subframe_len")
        if (self.frame_len - self.subframe_len) % self.subhop != 0:
            raise ValueError("(frame_len - subframe_len) must be divisible by subhop to yield integer number of subframes")
