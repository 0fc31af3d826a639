"""Batch orchestrator: the CALLER of the hot path.

Same public surface as the reference's audio_processing_framework.py
(``AudioProcessor`` :52-100, ``process_audio_batches_v2`` :580-894, alias ``process_audio_batches``
:898, ``restore_state_df_from_parquet`` :513-572) with one structural change: the reference hands
each file to ``proc.run`` separately (:183-207, serially or through a ProcessPoolExecutor
:249-290), which cannot feed a GPU.  Here a batch of files is handed to ``proc.run_batch`` in ONE
call when the processor has that hook (the GPU processors of this package do); processors
without it are still called per file through ``run``.  Row / state / parquet / attrs conventions
are the reference's, so downstream code (postprocess/*.py) sees the same DataFrames.

``debug_params["parallel"]`` selected a CPU process pool in the reference; the GPU path has no use
for it (one process per GPU) and it is accepted and ignored for ``run_batch`` processors.
Sharding a corpus over several GPUs is done one level up: every rank calls this function with
its own slice of the keys (see ``shard_keys``) and the per-clip rows are gathered afterwards
(parallel.gather_clip_stats).
"""
from __future__ import annotations

import dataclasses
import gc
import json
import time
from pathlib import Path
from typing import Any, Callable, Dict, List, Mapping, Optional, Protocol, Sequence, Tuple, runtime_checkable

import numpy as np
import pandas as pd


@runtime_checkable
class AudioProcessor(Protocol):
    """``name`` + ``run(audio_data, params) -> (results, state)``; optional ``setup(params)`` and
    ``run_batch(list_of_audio, params) -> list[(results, state)]``."""

    @property
    def name(self) -> str: ...

    def run(self, audio_data: np.ndarray, params: Dict[str, Any]) -> Tuple[Dict[str, Any], Dict[str, Any]]: ...


def _param_updates(obj: Any) -> Dict[str, Any]:
    if isinstance(obj, dict) and isinstance(obj.get("_param_updates"), dict):
        return obj["_param_updates"]
    return {}


def _namespaced(ns: str, d: Dict[str, Any]) -> Dict[str, Any]:
    return {f"{ns}__{k}": v for k, v in d.items()}


def _params_key(p: Dict[str, Any]) -> str:
    try:
        return json.dumps(p, sort_keys=True, default=repr)
    except Exception:
        return repr(sorted(p.items(), key=lambda kv: kv[0]))


def shard_keys(keys: Sequence[Any], rank: int, world: int) -> List[Any]:
    """Contiguous, near-equal slice of ``keys`` for ``rank`` of ``world`` (clips shard by file)."""
    n = len(keys)
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    return list(keys[lo:hi])


def _run_batch(*, dir_content, processors, params_global, params_by_processor, required_samples, rain_min_thr):
    """All files of one loaded batch through all processors.  Returns [{"row":..., "states":...}]."""
    items = []
    for file_key, meta in dir_content.items():
        audio = meta.get("file_contents")
        if audio is None:
            continue
        audio = np.asarray(audio)
        if audio.ndim != 1:
            raise ValueError(f"audio for {file_key} must be 1-D, got shape {audio.shape}")
        if audio.size < required_samples:
            continue
        row = {"file_key": file_key, "rain_actual": meta.get("raining", None)}
        if "synthetic_noise_info" in meta:
            row["synthetic_noise_info"] = meta["synthetic_noise_info"]
        items.append({"file_key": file_key, "meta": meta, "audio": audio, "row": row, "states": {},
                      "ctx": dict(params_global)})
    for proc in processors:
        over = params_by_processor.get(proc.name, {})
        # files whose chained parameters are identical go to the device together
        groups: Dict[str, List[int]] = {}
        pp: Dict[str, Dict[str, Any]] = {}
        for i, it in enumerate(items):
            p = dict(it["ctx"])
            p.update(over)
            k = _params_key(p)
            groups.setdefault(k, []).append(i)
            pp[k] = p
        for k, idxs in groups.items():
            p = pp[k]
            if hasattr(proc, "setup"):
                proc.setup(p)
            if hasattr(proc, "run_batch"):
                outs = proc.run_batch([items[i]["audio"] for i in idxs], p)
            else:
                outs = [proc.run(items[i]["audio"], p) for i in idxs]
            for i, (res, st) in zip(idxs, outs):
                it = items[i]
                res = dict(res) if isinstance(res, dict) else {"value": res}
                st = dict(st) if isinstance(st, dict) else {"state": st}
                st["file_key"] = it["file_key"]
                if "synthetic_noise_info" in it["meta"]:
                    st["synthetic_noise_info"] = it["meta"]["synthetic_noise_info"]
                it["states"][proc.name] = st
                it["row"].update(_namespaced(proc.name, res))
                upd = {}
                upd.update(_param_updates(res))
                upd.update(_param_updates(st))
                if upd:
                    it["ctx"].update(upd)
    out = []
    for it in items:
        row = it["row"]
        if "rain__rain_drops" in row and row["rain_actual"] is not None and rain_min_thr is not None:
            pred = bool(row["rain__rain_drops"] > rain_min_thr)
            row["rain__predicted"] = pred
            row["rain__mismatch"] = pred != bool(row["rain_actual"])
        out.append({"row": row, "states": it["states"]})
    return out


# -- parquet helpers ---------------------------------------------------------------------------
def _plain(value: Any) -> Any:
    """numpy-heavy nested values -> plain Python objects parquet can store."""
    if dataclasses.is_dataclass(value) and not isinstance(value, type):
        return _plain(dataclasses.asdict(value))
    if isinstance(value, np.ndarray):
        return value.tolist()
    if isinstance(value, np.generic):
        return value.item()
    if isinstance(value, type):
        return f"{value.__module__}.{value.__qualname__}"
    if isinstance(value, Mapping):
        return {k: _plain(v) for k, v in value.items()}
    if isinstance(value, (list, tuple)):
        return [_plain(v) for v in value]
    return value


_NMF = "normalized_mode_flux_by_mode"


def _state_rows_for_parquet(rows):
    safe = []
    for row in rows:
        r = {k: (_plain(v) if k != "features" else v) for k, v in row.items()}
        feats = r.get("features")
        if isinstance(feats, Mapping):
            feats = dict(feats)
            nmf = feats.pop(_NMF, None)
            if nmf is not None:
                nmf = np.asarray(nmf)
                if nmf.ndim != 2:
                    raise ValueError(f"features['{_NMF}'] must be 2-D when present; got shape {nmf.shape}")
                for m in range(nmf.shape[0]):
                    r[f"{_NMF}_{m}"] = nmf[m].tolist()
        if "features" in r:
            r["features"] = _plain(feats)
        safe.append(r)
    return safe


def _write_chunk(rows, path: Path) -> None:
    if not rows:
        return
    df = pd.DataFrame(rows)
    if "file_key" in df.columns and not df.empty:
        df = df.sort_values("file_key").reset_index(drop=True)
    df.to_parquet(path, index=False)


def _flush(results_rows, states_by_processor, save_dir: Path, prefix: str, idx: int):
    save_dir.mkdir(parents=True, exist_ok=True)
    res_paths, st_paths = [], {n: [] for n in states_by_processor}
    if results_rows:
        p = save_dir / f"{prefix}__results_part_{idx:05d}.parquet"
        _write_chunk(results_rows, p)
        res_paths.append(str(p))
    for name, rows in states_by_processor.items():
        if rows:
            p = save_dir / f"{prefix}__state__{name}_part_{idx:05d}.parquet"
            _write_chunk(_state_rows_for_parquet(rows), p)
            st_paths[name].append(str(p))
    return res_paths, st_paths


def restore_state_df_from_parquet(path) -> pd.DataFrame:
    """Inverse of the parquet-write transform: re-assembles ``normalized_mode_flux_by_mode_<i>``
    columns into ``features['normalized_mode_flux_by_mode']`` (2-D array) per row."""
    df = pd.read_parquet(path).copy()
    cols = sorted((c for c in df.columns if c.startswith(_NMF + "_")), key=lambda c: int(c.rsplit("_", 1)[1]))
    if not cols:
        return df
    feats = []
    for _, row in df.iterrows():
        f = dict(row["features"]) if isinstance(row.get("features"), dict) else {}
        parts = [row[c] for c in cols]
        if all(v is not None for v in parts):
            f[_NMF] = np.stack([np.asarray(v) for v in parts], axis=0)
        feats.append(f)
    df["features"] = feats
    return df.drop(columns=cols)


# -- orchestrator --------------------------------------------------------------------------------
def process_audio_batches_v2(
    *,
    processors: List[AudioProcessor],
    params_global: Dict[str, Any],
    params_by_processor: Optional[Dict[str, Dict[str, Any]]] = None,
    debug_params: Optional[Dict[str, Any]] = None,
    InputType: Optional[str] = None,
    test_vector_path: Optional[str] = None,
    query: Optional[str] = None,
    adse_engine=None,
    batch_size: int = 1000,
    max_files: Optional[int] = None,
    max_batch_save: int = 10_000,
    batch_save_dir: Optional[str] = "./save_dir",
    batch_save_prefix: str = "audio_processing_dump",
    local_cache: Optional[str] = None,
    localStatus: bool = True,
    get_keys_fn: Optional[Callable[..., List[Dict[str, Any]]]] = None,
    get_input_data_fn: Optional[Callable[..., Dict[str, Dict[str, Any]]]] = None,
    get_input_data_kwargs: Optional[Dict[str, Any]] = None,
) -> Tuple[pd.DataFrame, Dict[str, pd.DataFrame]]:
    """Run processors over a corpus in batches (arguments as in the reference, :580-707).

    ``get_keys_fn`` / ``get_input_data_fn`` must be supplied: key discovery and file / S3 / DB
    loading are host I/O outside this package (the reference's defaults live in audio_io.py and
    need boto3 / sqlalchemy / kaitaistruct).  The loader returns
    ``{file_key: {"file_contents": 1-D ndarray (float32 or int16), "raining": bool, ...}}``.
    """
    t_start = time.perf_counter()
    params_by_processor = params_by_processor or {}
    debug_params = debug_params or {}
    get_input_data_kwargs = get_input_data_kwargs or {}
    if max_batch_save is None:
        max_batch_save = 10_000
    if batch_save_dir is not None and max_batch_save <= 0:
        raise ValueError("max_batch_save must be > 0 when batch_save_dir is provided")
    save_dir = Path(batch_save_dir) if batch_save_dir is not None else None
    if "sample_rate" not in params_global or "check_duration" not in params_global:
        raise KeyError("params_global must contain 'sample_rate' and 'check_duration'.")
    if get_keys_fn is None or get_input_data_fn is None:
        raise ValueError("get_keys_fn and get_input_data_fn are required: the default loaders of the reference "
                         "(audio_io.get_keys / get_input_data) are host I/O outside this package")
    Fs = params_global["sample_rate"]
    check_duration = params_global["check_duration"]
    required_samples = int(Fs * check_duration)

    keys = get_keys_fn(InputType, test_vector_path=test_vector_path, query=query, adse_engine=adse_engine,
                       batch_size=batch_size, localStatus=localStatus)
    if max_files is not None:
        if max_files < 0:
            raise ValueError("max_files must be >= 0 or None")
        keys = keys[:max_files]
    print(f"received {len(keys)} test vectors" + ("" if max_files is None else f" (limited by max_files={max_files})"))

    results_rows: List[Dict[str, Any]] = []
    states_by_processor: Dict[str, List[Dict[str, Any]]] = {p.name: [] for p in processors}
    saved_res: List[str] = []
    saved_st: Dict[str, List[str]] = {p.name: [] for p in processors}
    flush_idx = 0
    print_mismatched = bool(debug_params.get("print_mismatched", False))
    debug_all = bool(debug_params.get("debug_all", False))
    rain_min_thr = debug_params.get("rain_drop_min_thr", params_global.get("rain_drop_min_thr"))
    total_batches = (len(keys) + batch_size - 1) // batch_size if batch_size > 0 else 1

    def flush():
        nonlocal flush_idx
        flush_idx += 1
        rp, sp = _flush(results_rows, states_by_processor, save_dir, batch_save_prefix, flush_idx)
        saved_res.extend(rp)
        for n, paths in sp.items():
            saved_st[n].extend(paths)

    for batch_idx, start in enumerate(range(0, len(keys), batch_size), start=1):
        print(f"Processing batch {batch_idx} of ~{total_batches}")
        dir_content = get_input_data_fn(keys[start:start + batch_size], InputType, Fs, check_duration, localStatus,
                                        local_cache, read_size=None, bytes_per_sample=2, **get_input_data_kwargs)
        outputs = _run_batch(dir_content=dir_content, processors=processors, params_global=params_global,
                             params_by_processor=params_by_processor, required_samples=required_samples,
                             rain_min_thr=rain_min_thr)
        for item in outputs:
            row = item["row"]
            if "rain__mismatch" in row and ((print_mismatched and row["rain__mismatch"]) or debug_all):
                rd = row.get("rain__rain_drop_count", row.get("rain__rain_drops"))
                print(f"[mismatch] {row['file_key']}  actual={row.get('rain_actual')}  "
                      f"predicted={row.get('rain__predicted')}  rain_drops={rd}")
            results_rows.append(row)
            for name, st in item["states"].items():
                states_by_processor[name].append(st)
        if save_dir is not None and max_batch_save > 0 and len(results_rows) >= max_batch_save:
            flush()
            results_rows.clear()
            for rows in states_by_processor.values():
                rows.clear()
        del dir_content
        gc.collect()

    if save_dir is not None and (results_rows or any(states_by_processor.values())):
        flush()

    results_df = pd.DataFrame(results_rows)
    if not results_df.empty:
        results_df = results_df.sort_values("file_key").reset_index(drop=True)
    results_df.attrs["saved_parquet_files"] = saved_res
    states_df: Dict[str, pd.DataFrame] = {}
    for name, rows in states_by_processor.items():
        df = pd.DataFrame(rows).sort_values("file_key").reset_index(drop=True) if rows else pd.DataFrame()
        df.attrs["saved_parquet_files"] = saved_st.get(name, [])
        states_df[name] = df
    wall = time.perf_counter() - t_start
    fps = (len(keys) / wall) if wall > 0 else None
    for df in [results_df, *states_df.values()]:
        df.attrs["wall_time_sec"] = wall
        df.attrs["num_files_processed_total"] = len(keys)
        df.attrs["files_per_sec_total"] = fps
    print(f"Total wall time: {wall:.3f} s")
    print(f"Total files processed: {len(keys)}")
    if fps is not None:
        print(f"Throughput: {fps:.3f} files/s")
    return results_df, states_df


process_audio_batches = process_audio_batches_v2
