"""Host logic of the orchestrator and the configuration layer (CPU only, fake processors)."""
import numpy as np
import pandas as pd
import pytest

from audio_processing_tools_b200.audio_processing_framework import (
    process_audio_batches, process_audio_batches_v2, restore_state_df_from_parquet, shard_keys)
from audio_processing_tools_b200.config import DetectorView, ResolvedParams, build_noise_config
from audio_processing_tools_b200.processors import BaseProcessor, RainProcessor, has_processor
from audio_processing_tools_b200.synth import MODES, default_params

FS = 100


class BatchProc:
    """run_batch-capable fake: records how the orchestrator grouped the files."""
    name = "energy"

    def __init__(self):
        self.calls = []

    def run_batch(self, audio_list, params):
        self.calls.append((len(audio_list), params.get("gain", 1.0)))
        out = []
        for a in audio_list:
            e = float(np.sum(np.asarray(a, np.float64) ** 2)) * params.get("gain", 1.0)
            out.append(({"energy": e}, {"n": a.size, "features": {"normalized_mode_flux_by_mode": np.ones((2, 3)), "decim": 1},
                                        "_param_updates": {"seen_energy": e > 5}}))
        return out

    def run(self, audio, params):
        return self.run_batch([audio], params)[0]


class PerFileProc:
    name = "rain"

    def run(self, audio, params):
        return {"rain_drops": int(params.get("seen_energy", False)) * 10}, {"chained": params.get("seen_energy")}


def _keys(n):
    def get_keys(InputType, **kw):
        return [{"source_file": f"f{i:03d}", "raining": bool(i % 2)} for i in range(n)]
    return get_keys


def _loader(keys, InputType, Fs, check_duration, localStatus, local_cache, read_size=None, bytes_per_sample=2, **kw):
    out = {}
    for k in keys:
        i = int(k["source_file"][1:])
        n = Fs * check_duration if i != 3 else 5            # f003 is too short and must be dropped
        out[k["source_file"]] = {"file_contents": np.full(n, 0.1 * (i + 1), np.float32), "raining": k["raining"]}
    return out


def test_orchestrator_batches_groups_and_chains(tmp_path):
    bp, fp = BatchProc(), PerFileProc()
    res, states = process_audio_batches_v2(
        processors=[bp, fp], params_global={"sample_rate": FS, "check_duration": 1},
        params_by_processor={"energy": {"gain": 2.0}}, debug_params={"rain_drop_min_thr": 3},
        batch_size=4, batch_save_dir=str(tmp_path), max_batch_save=5,
        get_keys_fn=_keys(10), get_input_data_fn=_loader)
    # 10 keys, one too short, batches of 4 -> run_batch called once per batch with the valid files
    assert [c[0] for c in bp.calls] == [3, 4, 2] and all(c[1] == 2.0 for c in bp.calls)
    saved = res.attrs["saved_parquet_files"]
    assert len(saved) == 2 and res.attrs["num_files_processed_total"] == 10
    full = pd.concat([pd.read_parquet(p) for p in saved]).sort_values("file_key").reset_index(drop=True)
    assert list(full["file_key"]) == [f"f{i:03d}" for i in range(10) if i != 3]
    assert {"energy__energy", "rain__rain_drops", "rain__predicted", "rain__mismatch", "rain_actual"} <= set(full.columns)
    # parameter chaining: energy's _param_updates reach the next processor of the same file
    e = full.set_index("file_key")
    for k, row in e.iterrows():
        assert row["rain__rain_drops"] == (10 if row["energy__energy"] > 5 else 0)
        assert row["rain__predicted"] == (row["rain__rain_drops"] > 3)
    st = restore_state_df_from_parquet(states["energy"].attrs["saved_parquet_files"][0])
    assert np.asarray(st["features"].iloc[0]["normalized_mode_flux_by_mode"]).shape == (2, 3)
    assert process_audio_batches is process_audio_batches_v2


def test_orchestrator_argument_errors():
    with pytest.raises(KeyError):
        process_audio_batches_v2(processors=[], params_global={"sample_rate": 1}, get_keys_fn=_keys(1), get_input_data_fn=_loader)
    with pytest.raises(ValueError):
        process_audio_batches_v2(processors=[], params_global={"sample_rate": 1, "check_duration": 1},
                                 get_keys_fn=_keys(1), get_input_data_fn=_loader, max_files=-1, batch_save_dir=None)
    with pytest.raises(ValueError):
        process_audio_batches_v2(processors=[], params_global={"sample_rate": 1, "check_duration": 1})


def test_shard_keys_partition():
    keys = list(range(1003))
    parts = [shard_keys(keys, r, 8) for r in range(8)]
    assert sum(parts, []) == keys and max(map(len, parts)) - min(map(len, parts)) <= 1


def test_processor_adapters():
    def fn(audio, **params):
        return 7, 0.5, {"rain_drop_count": 7, "other": 1}
    p = RainProcessor(name="rain", fn=fn)
    res, st = p.run(np.zeros(FS, np.float32), {"sample_rate": FS, "check_duration": 1})
    assert res["rain_drops"] == 7 and res["rain_drop_count"] == 7 and st["processor"] == "rain" and "latency_s" in res
    with pytest.raises(TypeError):
        p.run([0.0] * FS, {})
    with pytest.raises(ValueError):
        p.run(np.zeros((2, FS)), {})
    with pytest.raises(ValueError):
        p.run(np.zeros(FS - 1), {"sample_rate": FS, "check_duration": 1})
    assert has_processor([p], "rain") and not has_processor([p], "noise")
    assert isinstance(p, BaseProcessor)


def test_config_precedence_and_resolution():
    params = default_params(q=0.3, suppressor={"q": 0.1, "win_sec": 0.25}, fmin=500.0, fmax=3000.0)
    params["detector"]["td_gate_threshold"] = 3.0
    cfg = build_noise_config(11162, params)
    assert cfg.q == 0.3 and cfg.win_sec == 0.25              # flat > suppressor > default
    assert cfg.operating_band == (500.0, 3000.0)             # legacy fmin/fmax
    assert DetectorView(cfg).get("td_gate_threshold") == 3.0
    rp = ResolvedParams(build_noise_config(11162, default_params()), 11162)
    c = rp.c
    assert (c.band_lo, c.band_hi, rp.K, rp.F, rp.M) == (10, 80, 71, 129, 5)
    assert [(c.mode_lo[i], c.mode_hi[i]) for i in range(5)] == [(11, 14), (19, 24), (35, 41), (54, 58), (73, 76)]
    assert c.warmup_need == 21 and c.n_sos == 2 and c.padlen == 15
    assert c.trk_eta == pytest.approx(2.0 / 44.0) and c.bl_eta == pytest.approx(2.0 / 45.0)
    assert len(MODES) == 5
