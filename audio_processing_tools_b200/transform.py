"""Entry points of the reference's transform.py for the DSD path (the reference module cannot even be imported
as shipped: transform.py:25 imports a module path that does not exist).

Kept with the same names and meaning: ``emulator_output_to_df`` (:51-68), ``reverse_binning_func`` / ``dsd_weights`` /
``add_weighted_dsd_data`` (:123-145), ``process_audio_file_dsd`` (:251-313).  Not carried over: the two scipy
one-liners ``butter_bandpass_filter`` / ``get_real_fft_df`` (:29-48; plotting helpers that nothing on the DSD path
calls -- this package computes nothing on the CPU) and ``dsd_from_audio_keys`` (:316-403; S3 fetch + Postgres upsert).  The compute of
``process_audio_file_dsd`` runs on the GPU (host_analysis.device_dsd_processing_emulator); fetching from S3,
Mark-3 container parsing and the database upsert (``dsd_from_audio_keys`` :316-403) are host I/O outside this
package, so ``process_audio_file_dsd`` takes the decoded int16 signal and its metadata instead of an S3 key,
and ``process_audio_signals_dsd`` does the same for a batch.
"""
from __future__ import annotations

import datetime as dt
from typing import Any, Dict, Sequence

import numpy as np
import pandas as pd

from .host_analysis.device_dsd_processing_emulator import DsdProcessingEmualtor

RAIN_ENERGY_THRESHOLD = 0.6
RAIN_LOG_FACTOR = 0.6


def emulator_output_to_df(output, device_id, audio_start_timestamp, output_interval_min=1):
    cols = [f"dsd{i}" for i in range(32)] + [f"pft{i}" for i in range(30)] + [f"fft{i}" for i in range(38)]
    df = pd.DataFrame(output, columns=cols)
    df["time"] = pd.date_range(audio_start_timestamp + dt.timedelta(minutes=1), periods=len(df),
                               freq=f"{output_interval_min}min")
    df["device"] = device_id
    return df


def reverse_binning_func(drop_bin, threshold=RAIN_ENERGY_THRESHOLD):
    return (((np.e ** (drop_bin * np.log(1.13))) - 1) / RAIN_LOG_FACTOR) + threshold


dsd_weights = {f"dsd{i}": reverse_binning_func(i) for i in range(32)}


def add_weighted_dsd_data(df, weights=dsd_weights.values(), add_to_df=True, add_weighted_dsd_sum=False):
    weighted = (df[[f"dsd{i}" for i in range(32)]] * weights).add_suffix("_weighted")
    if add_weighted_dsd_sum:
        weighted["weighted_dsd_sum"] = weighted.sum(axis=1)
    return pd.concat([df, weighted], axis=1) if add_to_df else weighted


def process_audio_signals_dsd(signals: Sequence[np.ndarray], metadata: Sequence[Dict[str, Any]], keys: Sequence[str],
                              verbose=False, device=0):
    """Batch form of process_audio_file_dsd: first 60 s of every int16 signal through the emulator in one GPU pass."""
    if not signals:
        return []
    rates = {int(m["sample_rate"]) for m in metadata}
    if len(rates) != 1:
        raise ValueError("all signals of a batch must share one sample rate")
    fs = rates.pop()
    clipped = []
    for sig in signals:
        sig = np.asarray(sig)
        clipped.append(sig[: 60 * fs] if round(len(sig) / fs) > 60 else sig)
    model = DsdProcessingEmualtor(fs=fs, frame_length=512, hop_length=512, bwindow=False, ts=0, verbose=verbose, device=device)
    outs = model.process_audio_batch(clipped, [0] * len(clipped))
    dfs = []
    for out, meta, key in zip(outs, metadata, keys):
        df = emulator_output_to_df(out, meta["device_id"], meta["time"])
        df["key"] = key
        dfs.append(add_weighted_dsd_data(df, add_weighted_dsd_sum=True))
    return dfs


def process_audio_file_dsd(key, sig, metadata, verbose=False, reprocess=False, device=0):
    """transform.process_audio_file_dsd (:251-313) for an already fetched and parsed Mark-3 file."""
    return process_audio_signals_dsd([sig], [metadata], [key], verbose=verbose, device=device)[0]
