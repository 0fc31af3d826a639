#!/usr/bin/env python
"""Benchmark of the hot path (contract: prints ONE JSON line on rank 0).

Workload (BASELINE.json configs[2] at N = 1, configs[3] at N > 1): the full pipeline (STFT -> noise floor ->
rain-event detection) on 1,000 x 10-min synthetic clips.  A "step" is one pass of the hot path over the batch.
  * --scaling strong (default; BASELINE configs[3] as written): the SAME --clips batch is split over the ranks
    in contiguous index ranges (parallel.shard_range: 500 / 250 / 125 clips per GPU), no data-path collective,
    one NCCL all-gather of the per-clip statistics rows at the end of each step.
  * --scaling weak: --clips per GPU (independent replicas + the same gather).
`value` is device-timed with the int16 PCM already resident in HBM.  `e2e` is the same metric through the
reference-facing plugin call -- RainDetectorProcessor.run_batch on a list of host arrays, default flags -- with the
staging, host<->device copies and the result dictionaries inside the timed region; `e2e_abi` is the C-ABI host entry
point on one pinned buffer (no Python packaging), kept beside it.

Other driver-visible workloads: --workload features_1h (BASELINE configs[1]: one 1-hour clip, STFT + band energies)
and --workload sweep --n-fft N --hop H (configs[4]: features stage at another frame size).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scaling strong|weak] [--workload ...]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

FS = 11162
HOP = 128
FFT_MODE = {"f64": True, "f32": False, "tc": "tc"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--clips", type=int, default=1000, help="clips of the batch (strong: in total; weak: per GPU)")
    ap.add_argument("--clip-seconds", type=float, default=600.0)
    ap.add_argument("--base-clips", type=int, default=16, help="distinct synthetic clips tiled to --clips")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--fft", default="f64", choices=("f64", "f32", "tc"),
                    help="STFT arithmetic: f64 = the reference's (default, bit-exact events), f32 = float32 FFT, tc = tensor-core DFT (both tolerance paths)")
    ap.add_argument("--scaling", default="strong", choices=("strong", "weak"))
    ap.add_argument("--workload", default="full", choices=("full", "features_1h", "sweep", "bne", "roe", "dsd"),
                    help="full / features_1h / sweep: the hot path; bne, roe, dsd: the engines beside it (SURVEY 8(f) rows 1-3)")
    ap.add_argument("--config", type=int, default=None, choices=(1, 2, 3, 4),
                    help="shorthand for BASELINE configs[i]: 1 = --workload features_1h, 2 / 3 = --workload full (3 under torchrun), "
                         "4 = --workload sweep at --n-fft / --hop (default 1024 / 256)")
    ap.add_argument("--n-fft", type=int, default=256)
    ap.add_argument("--hop", type=int, default=128)
    ap.add_argument("--write-spectra", action="store_true", help="features workloads: also write S (the HBM-bound variant)")
    ap.add_argument("--e2e-input", default="int16", choices=("int16", "float32"))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--operating-band", type=float, nargs=2, default=None,
                    help="experiment knob: another operating band (the BASELINE workload uses the default 400..3500 Hz = 71 bins)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-cpu-reference", action="store_true", help="skip the leg that times the unmodified reference")
    args = ap.parse_args()
    if args.config is not None:
        if args.config == 1:
            args.workload = "features_1h"
        elif args.config in (2, 3):
            args.workload = "full"
        else:
            args.workload = "sweep"
            if (args.n_fft, args.hop) == (256, 128):
                args.n_fft, args.hop = 1024, 256
    return args


def make_base_clips(n_base, seconds, seed0):
    from audio_processing_tools_b200.synth import batch_clip_spec, synth_clip_i16
    clips = []
    for i in range(n_base):
        seed, lam = batch_clip_spec(seed0 + i)
        clips.append(synth_clip_i16(seconds, seed, lam))
    return clips


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------------------
# CPU legs (the only code of this file that touches oracle/)
# ------------------------------------------------------------------------------------------------------------
def cpu_port(params, seconds, n_threads, target_clips=None, target_wall=12.0):
    """Times the CPU oracle (C port of the reference's algorithm) on a bounded sample of the workload:
    a short calibration pass sizes the sample to about `target_wall` seconds of wall time on all threads."""
    from oracle import oracle
    from audio_processing_tools_b200.synth import batch_clip_spec, synth_clip_i16
    oracle.build()
    distinct = [synth_clip_i16(seconds, *batch_clip_spec(900 + i)) for i in range(4)]

    def run(n):
        clips = [distinct[i % len(distinct)] for i in range(n)]
        t0 = time.perf_counter()
        oracle.process_batch_i16(clips, params, n_threads=n_threads)
        return time.perf_counter() - t0

    if target_clips is None:
        n0 = max(2, 2 * n_threads)
        dt0 = run(n0)
        target_clips = int(min(4096, max(n0, round(n0 * target_wall / max(dt0, 1e-3)))))
        target_clips = max(n_threads, target_clips // n_threads * n_threads)
    dt = run(target_clips)
    return target_clips * seconds / dt, target_clips, dt


_REF = {}


def reference_available():
    """The unmodified reference (baseline/_ref: `pip install --target` of /root/reference; or /root/reference itself)
    behind oracle/refharness (librosa stand-in + import stubs for its absent third-party modules)."""
    if "ok" in _REF:
        return _REF["ok"]
    try:
        sys.path.insert(0, os.path.join(REPO, "oracle"))
        import refharness
        _REF["root"] = refharness.install()
        from audio_processing_tools.audio_processing_framework import process_audio_batches_v2  # noqa: F401
        from audio_processing_tools.edge.rain_signal_processor import RainDetectorProcessor  # noqa: F401
        _REF["ok"] = True
    except Exception as exc:  # pragma: no cover
        _REF["ok"] = False
        _REF["why"] = f"{type(exc).__name__}: {exc}"
    return _REF["ok"]


def cpu_reference(seconds, n_clips, workers, geom=None):
    """The reference's own CPU path: process_audio_batches_v2 with its RainDetectorProcessor, injected
    synthetic loaders, parallel=True over `workers` processes (the reference's rule: cpu_count - 1)."""
    import contextlib
    import io
    from audio_processing_tools.audio_processing_framework import process_audio_batches_v2
    from audio_processing_tools.edge.rain_signal_processor import RainDetectorProcessor
    from audio_processing_tools_b200.synth import MODES, batch_clip_spec, pcm_to_f32, synth_clip_i16
    distinct = [pcm_to_f32(synth_clip_i16(seconds, *batch_clip_spec(900 + i))) for i in range(4)]
    names = [f"clip{i:05d}" for i in range(n_clips)]

    def get_keys(InputType, **kw):
        return [{"source_file": n, "raining": False} for n in names]

    def loader(keys, InputType, Fs, check_duration, localStatus, local_cache, read_size=None, bytes_per_sample=2, **kw):
        return {k["source_file"]: {"file_contents": distinct[int(k["source_file"][4:]) % 4], "raining": False} for k in keys}

    params = {"sample_rate": FS, "check_duration": seconds, "detector": {"mode_bands": [tuple(m) for m in MODES]}, **(geom or {})}
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        res, _ = process_audio_batches_v2(processors=[RainDetectorProcessor()], params_global=params,
                                          debug_params={"parallel": workers > 1, "num_workers": workers},
                                          batch_size=n_clips, batch_save_dir=None,
                                          get_keys_fn=get_keys, get_input_data_fn=loader)
    dt = time.perf_counter() - t0
    assert len(res) == n_clips
    return n_clips * seconds / dt, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation on the host cores.  The unmodified reference when it is
    importable (baseline/_ref), on a bounded sample; the C port of its algorithm is timed beside it."""
    if rank != 0:
        return
    from audio_processing_tools_b200.synth import default_params
    geom = {} if (args.n_fft, args.hop) == (256, 128) else {"n_fft": args.n_fft, "hop": args.hop}
    params = default_params(check_duration=args.clip_seconds, **geom,
                            **({"operating_band": tuple(args.operating_band)} if args.operating_band else {}))
    cores = os.cpu_count() or 1
    workers = max(1, cores - 1)
    port_v, port_n, port_dt = cpu_port(params, args.clip_seconds, cores, target_wall=8.0)
    port = {"value": port_v, "unit": "audio-s/s", "cores": cores, "kind": "port",
            "sample": f"{port_n} x {args.clip_seconds:g}s clips, C port of the reference algorithm (oracle/apt_oracle.c), {cores} threads, {port_dt:.1f}s"}
    common = {"impl": "reference", "metric": "audio_seconds_per_second", "unit": "audio-s/s", "n_gpus": args.gpus,
              "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": args.scaling,
              "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic"}
    if reference_available():
        # bounded sample: `workers` clips of 60 s per step (the reference needs ~4 s per 60 s of audio per core)
        sec, n = 60.0, workers
        for _ in range(min(1, args.warmup)):
            cpu_reference(sec, n, workers, geom)
        vals, times = [], []
        for _ in range(max(1, min(args.steps, 5))):
            v, dt = cpu_reference(sec, n, workers, geom)
            vals.append(v)
            times.append(dt)
        value = float(np.mean(vals))
        sample = (f"{n} x {sec:g}s clips per step (bounded sample of the {args.clips} x {args.clip_seconds:g}s batch), the UNMODIFIED reference "
                  f"(baseline/_ref) through its own process_audio_batches_v2 + RainDetectorProcessor, parallel=True, {workers} worker processes "
                  f"(its own default rule cpu_count - 1); librosa 0.11 is not installable offline: oracle/refharness supplies the stft stand-in")
        print(json.dumps({**common, "value": value, "ms_per_step": float(np.mean(times) * 1e3),
                          "config": {"workload": f"full pipeline, {args.clips} x {args.clip_seconds:g}s synthetic clips (BASELINE configs[{4 if geom else 2}]) n_fft={args.n_fft} hop={args.hop}; "
                                                 f"the CPU arm runs a bounded sample per step"},
                          "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": workers, "kind": "reference", "sample": sample},
                          "cpu_port": port,
                          "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    vals, times = [], []
    for _ in range(max(1, args.steps)):
        v, n, dt = cpu_port(params, args.clip_seconds, cores, target_clips=port_n)
        vals.append(v)
        times.append(dt)
    value = float(np.mean(vals))
    port["value"] = value
    print(json.dumps({**common, "value": value, "ms_per_step": float(np.mean(times) * 1e3),
                      "config": {"workload": f"full pipeline, {args.clips} x {args.clip_seconds:g}s synthetic clips (BASELINE configs[{4 if geom else 2}]) n_fft={args.n_fft} hop={args.hop}; "
                                             f"the CPU arm runs a bounded sample per step"},
                      "cpu_baseline": port, "reference_unavailable": _REF.get("why"),
                      "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def load_peak():
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    for rel in (("profiles", "r2", "kernel_traffic_r2.json"), ("profiles", "r1", "kernel_traffic_r1.json")):
        try:
            return json.load(open(os.path.join(REPO, *rel))), "/".join(rel)
        except Exception:
            continue
    return None, None


def timed_steps(torch, dist, world, dev, step, steps, warmup, local_rank):
    """W warm-up steps, then `steps` steps between CUDA events with a barrier + synchronize on both sides; max over ranks."""
    for _ in range(max(3, warmup)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    return ms_total / steps, sampler.summary()


def run_neighbour(args, torch, local_rank, dev):
    """band noise estimator, legacy RoE detector, DSD emulator.  `value`: the C-ABI call timed with
    CUDA events, inputs resident in HBM; `e2e`: the public Python call on host arrays (H2D, kernels, D2H, packaging);
    `cpu_baseline`: the unmodified reference (baseline/_ref) when importable, else its numpy restatement (oracle/), one core."""
    from audio_processing_tools_b200 import _lib
    from audio_processing_tools_b200.synth import batch_clip_spec, pcm_to_f32, synth_clip_i16
    kind = args.workload
    peak, peak_src = load_peak()
    if kind == "bne":
        from audio_processing_tools_b200.edge.band_noise_processor import BandNoiseEstimatorProcessor
        n, sec = 512, 60.0
        base = [synth_clip_i16(sec, *batch_clip_spec(i)) for i in range(8)]
        batch = [base[i % 8] for i in range(n)]
        proc = BandNoiseEstimatorProcessor()
        call = lambda: proc.run_batch(batch, {"sample_rate": FS})
        N = 512                       # BandNoiseEstimatorConfig.frame_len, the reference's default
        frames = n * (int(FS * sec) // N)
        bytes_algo = n * int(FS * sec) * 2 + frames * (_lib.BNE_FRAME_F * 8 + 1 + _lib.BNE_MAX_S * 8)
        what = f"band noise estimator (edge/band_noise_estimator.py), {n} x {sec:g}s clips, frame {N} (the reference default)"

        def cpu(nc):
            if reference_available():
                from audio_processing_tools.edge.band_noise_processor import BandNoiseEstimatorProcessor as Ref
                r = Ref()
                t0 = time.perf_counter()
                for i in range(nc):
                    r.run(pcm_to_f32(base[i % 8]), {"sample_rate": FS})
                return time.perf_counter() - t0, "reference"
            from oracle import band_noise_oracle
            t0 = time.perf_counter()
            for i in range(nc):
                band_noise_oracle.run(pcm_to_f32(base[i % 8]), {"sample_rate": FS})
            return time.perf_counter() - t0, "port"
        cpu_n = 32
    elif kind == "roe":
        from audio_processing_tools_b200.edge import dsp_rain_detection as roe
        n, sec = 2000, 10.0
        base = [synth_clip_i16(sec, *batch_clip_spec(i)) for i in range(8)]
        batch = [base[i % 8] for i in range(n)]
        call = lambda: roe.rain_detection_algo_batch(batch, **roe.default_params)
        frames = n * (1 + int(FS * sec) // 128)
        bytes_algo = n * int(FS * sec) * 2 + frames * _lib.ROE_FRAME_F * 8
        what = f"legacy RoE detector (edge/dsp_rain_detection.py rain_detection_algo), {n} x {sec:g}s clips"

        def cpu(nc):
            if reference_available():
                import audio_processing_tools.edge.dsp_rain_detection as ref
                import contextlib
                import io
                t0 = time.perf_counter()
                with contextlib.redirect_stdout(io.StringIO()):
                    for i in range(nc):
                        ref.rain_detection_algo(pcm_to_f32(base[i % 8]), **roe.default_params)
                return time.perf_counter() - t0, "reference"
            from oracle import roe_oracle
            t0 = time.perf_counter()
            for i in range(nc):
                roe_oracle.rain_detection_algo(pcm_to_f32(base[i % 8]), **roe.default_params)
            return time.perf_counter() - t0, "port"
        cpu_n = 32
    else:
        from audio_processing_tools_b200.host_analysis.device_dsd_processing_emulator import DsdProcessingEmualtor
        n, sec = 256, 300.0
        base = [synth_clip_i16(sec, *batch_clip_spec(i)) for i in range(4)]
        batch = [base[i % 4] for i in range(n)]
        em = DsdProcessingEmualtor(fs=FS, frame_length=512, hop_length=512, bwindow=False, ts=0)
        call = lambda: em.process_audio_batch(batch, [0.0] * n)
        frames = n * (int(FS * sec) // 512)
        bytes_algo = n * int(FS * sec) * 2 + n * int(np.ceil(sec / 60.0)) * 100 * 8
        what = f"DSD emulator (host_analysis/device_dsd_processing_emulator.py), {n} x {sec:g}s clips, frame 512"

        def cpu(nc):
            if reference_available():
                from audio_processing_tools.host_analysis.device_dsd_processing_emulator import DsdProcessingEmualtor as Ref
                t0 = time.perf_counter()
                for i in range(nc):
                    Ref(fs=FS, frame_length=512, hop_length=512, bwindow=False, ts=0, verbose=False).process_audio_data(base[i % 4], 0)
                return time.perf_counter() - t0, "reference"
            from oracle import dsd_oracle
            t0 = time.perf_counter()
            for i in range(nc):
                dsd_oracle.process_audio_data(base[i % 4], 0.0)
            return time.perf_counter() - t0, "port"
        cpu_n = 64
    audio_s = n * sec
    _lib.DEVICE_TIMING = True
    for _ in range(max(3, args.warmup)):
        call()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    samp = ClockSampler(local_rank)
    samp.start()
    dev_ms, wall = [], []
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        call()
        wall.append(time.perf_counter() - t0)
        dev_ms.append(_lib.LAST_DEVICE_MS[kind])
    samp.stop_flag.set()
    samp.join(timeout=2)
    ms = float(np.mean(dev_ms))
    res = {"metric": "audio_seconds_per_second", "value": audio_s / (ms * 1e-3), "unit": "audio-s/s", "n_gpus": 1, "steps": args.steps,
           "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": what + " (SURVEY 8(f)); a 256 MB write flushes L2 before every timed call", "clips": n, "clip_seconds": sec},
           "roofline": {"bound": "hbm", "kernel": f"apt_{kind}_run (all its kernels + table uploads, C-ABI call timed with CUDA events)",
                        "achieved": bytes_algo / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": bytes_algo / (ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes": bytes_algo},
           "e2e": {"value": audio_s / float(np.mean(wall)), "unit": "audio-s/s", "ms_per_step": float(np.mean(wall) * 1e3),
                   "h2d_bytes_per_step": int(n * int(FS * sec) * 2), "d2h_bytes_per_step": int(bytes_algo - n * int(FS * sec) * 2),
                   "note": "public Python call on a list of host arrays: concatenate, H2D, kernels, D2H, per-clip result packaging; wall clock"},
           "clocks": samp.summary(), "gpu_launches": None}
    if not args.no_cpu:
        dt, how = cpu(cpu_n)
        res["cpu_baseline"] = {"value": cpu_n * sec / dt, "unit": "audio-s/s", "cores": 1, "kind": how,
                               "sample": f"{cpu_n} x {sec:g}s clips, one process ({dt:.1f}s)"}
    print(json.dumps(res), flush=True)


def run_features(args, torch, dist, rank, world, local_rank, dev):
    """BASELINE configs[1] / configs[4]: features stage (framing, window, FFT, power, band energies) on one long clip."""
    from audio_processing_tools_b200 import _lib
    from audio_processing_tools_b200.config import build_noise_config
    from audio_processing_tools_b200.engine import BatchEngine
    from audio_processing_tools_b200.synth import default_params, synth_clip_i16
    import ctypes as C
    n_fft, hop = (256, 128) if args.workload == "features_1h" else (args.n_fft, args.hop)
    seconds = 3600.0
    params = default_params(check_duration=seconds, n_fft=n_fft, hop=hop)
    cfg = build_noise_config(FS, params)
    eng = BatchEngine(cfg, FS, device=local_rank, fft_f64=FFT_MODE[args.fft])
    base = synth_clip_i16(60.0, 7 + rank, 3.0)
    N = int(FS * seconds)
    pcm = np.tile(base, N // base.size + 1)[:N]
    n_clips = max(1, args.clips if args.clips != 1000 else 1)
    plan = eng.plan_for([N] * n_clips)
    pcm_dev = torch.from_numpy(pcm).to(dev).repeat(n_clips)
    want = ("band_energy",) + (("S",) if args.write_spectra else ())
    bufs = eng.alloc_outputs(plan, want, full=False)

    def step():
        eng.run_device(plan, pcm_dev, bufs, full=False)

    ms_step, clocks = timed_steps(torch, dist, world, dev, step, args.steps, args.warmup, local_rank)
    T = plan.nF // n_clips
    F = n_fft // 2 + 1
    bytes_algo = n_clips * (N * 2 + T * 24 + (T * F * 8 if args.write_spectra else 0))
    peak, peak_src = load_peak()
    ach = bytes_algo / (ms_step * 1e-3) / 1e9
    res = {"metric": "audio_seconds_per_second", "value": world * n_clips * seconds / (ms_step * 1e-3), "unit": "audio-s/s",
           "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": {"f64": "f32+f64", "f32": "f32", "tc": "f16x2 limbs, f32 accumulate"}[args.fft], "data": "synthetic",
           "config": {"workload": f"features stage (STFT + band energies{' + spectra' if args.write_spectra else ''}), {n_clips} x 1-hour clip per GPU, "
                                  f"n_fft={n_fft} hop={hop} (BASELINE configs[{1 if args.workload == 'features_1h' else 4}])",
                      "l2": "input %.0f MB per step; L2 flushed by the %s" % (plan.nS * 2 / 1e6, "spectra written" if args.write_spectra else "80 MB input itself (half of L2: partly resident)")},
           "roofline": {"bound": "hbm", "kernel": ("tcdft256_kernel" if args.fft == "tc" and (n_fft, hop) == (256, 128) else "stft256_kernel") if (n_fft == 256 and hop <= 128) else "stft_generic_kernel",
                        "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                        "algorithmic_bytes": bytes_algo, "peak_source": peak_src},
           "clocks": clocks, "gpu_launches": int(eng.last_launches * args.steps)}
    if rank == 0:
        print(json.dumps(res), flush=True)
    eng.close()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from audio_processing_tools_b200 import _lib
    from audio_processing_tools_b200.config import build_noise_config
    from audio_processing_tools_b200.engine import BatchEngine
    from audio_processing_tools_b200.parallel import bind_near_gpu, gather_clip_stats, shard_range
    from audio_processing_tools_b200.synth import default_params

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.workload in ("bne", "roe", "dsd"):
        if rank == 0:
            run_neighbour(args, torch, local_rank, dev)
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload != "full":
        run_features(args, torch, dist, rank, world, local_rank, dev)
        if world > 1:
            dist.destroy_process_group()
        return

    geom = {} if (args.n_fft, args.hop) == (256, 128) else {"n_fft": args.n_fft, "hop": args.hop}   # configs[4]: the whole pipeline at another frame size
    params = default_params(check_duration=args.clip_seconds, **geom,
                            **({"operating_band": tuple(args.operating_band)} if args.operating_band else {}))
    cfg = build_noise_config(FS, params)
    eng = BatchEngine(cfg, FS, device=local_rank, fft_f64=FFT_MODE[args.fft])
    N = int(FS * args.clip_seconds)
    strong = args.scaling == "strong"
    if strong:
        lo, hi = shard_range(args.clips, rank, world)     # contiguous index ranges of ONE batch (BASELINE configs[3])
        counts = [shard_range(args.clips, r, world)[1] - shard_range(args.clips, r, world)[0] for r in range(world)]
    else:
        lo, hi = rank * args.clips, (rank + 1) * args.clips
        counts = [args.clips] * world
    n_clips = hi - lo
    total_clips = sum(counts)
    n_base = max(1, min(args.base_clips, args.clips))
    # clip i of the batch is base clip i % n_base: every rank generates the same base set, so that the strong split
    # really is one batch cut in pieces
    base = make_base_clips(n_base, args.clip_seconds, seed0=0)
    plan = eng.plan_for([N] * n_clips)
    T = 1 + N // args.hop

    # device-resident shard (13.4 GB int16 at 1 000 clips, >> L2)
    base_dev = torch.from_numpy(np.stack(base)).to(dev)
    idx = torch.arange(lo, hi, device=dev) % n_base
    pcm_dev = base_dev[idx].contiguous().reshape(-1)
    del base_dev, idx
    bufs = eng.alloc_outputs(plan, (), full=True)
    gathered = {}

    def step():
        eng.run_device(plan, pcm_dev, bufs, full=True)
        if world > 1:
            gathered["rows"] = gather_clip_stats(bufs["clip_stats"], counts, clip_id_base=0)

    ms_step, clocks = timed_steps(torch, dist, world, dev, step, args.steps, args.warmup, local_rank)
    launches_per_step = eng.last_launches
    if world > 1:
        rows = gathered["rows"]
        assert rows.shape[0] == total_clips and bool((rows[:, 0] == torch.arange(total_clips, device=dev)).all()), \
            "gathered clip rows are not in global clip order"
    # per-kernel breakdown: a second, separate pass with the library's event marks on the launch stream (marks put the
    # whole pipeline on one stream in one time segment, so this pass is not the one that is reported as `value`)
    import ctypes as C
    eng.L.apt_plan_enable_timing(plan.h, 1)
    for _ in range(args.steps):
        eng.run_device(plan, pcm_dev, bufs, full=True)
    torch.cuda.synchronize()
    kms = (C.c_float * len(_lib.KERNEL_NAMES))()
    eng.L.apt_plan_kernel_ms(plan.h, kms)
    eng.L.apt_plan_enable_timing(plan.h, 0)
    audio_s = total_clips * args.clip_seconds
    value = audio_s / (ms_step * 1e-3)

    # rooflines: per kernel (device events inside the library, serial pass) and for the whole pipeline
    K = eng.rp.K
    nF = plan.nF
    n_mode_bins = sum(max(0, eng.rp.c.mode_band_hi[i] - eng.rp.c.mode_band_lo[i] + 1) for i in range(eng.rp.M))
    nls = max(8, (n_mode_bins + 7) // 8 * 8)
    bytes_algo = total_clips * (N * 2 + T * 9) + total_clips * 32                 # SURVEY 8(d) formula, int16 in
    kernel_bytes = {   # algorithmic bytes each kernel must move in this decomposition
        "stft256_kernel": plan.nS * 2 + nF * K * 4,
        "td_features_kernel": plan.nS * 2 + nF * 4,
        "trk1_kernel": nF * n_mode_bins * 4 + nF * nls * 4,
        "flux_kernel": nF * n_mode_bins * 4 + nF * nls * 4 + nF * 32,
        "base_kernel": 2 * nF * 32,
        "decide_kernels": nF * 32 + nF * 4 + nF * 9 + nF,
        "trk2_kernel": nF * K * 4 + nF + nF * K * 4,
        "db_kernel": nF * K * 4,
        "select_kernels": nF * K * 4,
        "finalize_kernel": n_clips * 64,
    }
    kms_step = {name: float(kms[i]) / args.steps for i, name in enumerate(_lib.KERNEL_NAMES)}
    dom = max(kms_step, key=lambda k: kms_step[k])
    peak, peak_src = load_peak()
    ach = kernel_bytes[dom] / (kms_step[dom] * 1e-3) / 1e9 if kms_step[dom] > 0 else 0.0
    traffic, traffic_note = None, None
    tj, tj_path = load_traffic()
    if tj is not None:
        try:
            key = {"decide_kernels": "decide_kernel", "select_kernels": "sel_collect_kernel", "db_kernel": "dbsum_kernel"}.get(dom, dom)
            traffic = tj["kernels"][key]["dram_bytes_per_frame"] * nF
            traffic_note = "ncu dram__bytes_read + write per frame on %s (%s), scaled by this launch's frames" % (tj["workload"], tj_path)
        except Exception:
            pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                "algorithmic_bytes": kernel_bytes[dom],
                "kernel_ms_per_step": kms_step, "kernel_ms_sum": sum(kms_step.values()),
                "kernel_gbs": {k: (kernel_bytes[k] / (v * 1e-3) / 1e9 if v > 0 else None) for k, v in kms_step.items()},
                "pipeline": {"bytes_algo": bytes_algo, "achieved_gbs": bytes_algo / (ms_step * 1e-3) / 1e9,
                             "frac": bytes_algo / (ms_step * 1e-3) / 1e9 / peak / world,
                             "flops_algo": 17000.0 * nF * world, "achieved_tflops": 17000.0 * T * total_clips / (ms_step * 1e-3) / 1e12,
                             "compute_frac_fp32_peak": 17000.0 * T * total_clips / (ms_step * 1e-3) / 74.4e12 / world},
                "note": "kernel_ms_per_step comes from a serial pass (one stream, one time segment, event marks); the timed steps run "
                        "the kernels pipelined over time segments on one stream per kernel kind, so ms_per_step is below kernel_ms_sum. "
                        "HBM fraction reported as the contract requires; at n_fft=256 the dominant kernels are bound by the FP64 / issue "
                        "pipes and the serial kernels by dependent-issue latency (DESIGN.md section 4)"}

    result = {
        "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None,
        "dtype": {"f64": "f32+f64", "f32": "f32", "tc": "f16 limbs (tensor cores) + f32 + f64"}[args.fft],
        "data": "synthetic",
        "config": {"workload": (f"full pipeline (STFT -> noise floor -> rain events), {total_clips} x {args.clip_seconds:g}s clips"
                                + (f" in one batch split over {world} GPU(s) ({', '.join(str(c) for c in sorted(set(counts)))} per GPU)" if strong
                                   else f" ({args.clips} per GPU)")
                                + f", fs=11162 n_fft={args.n_fft} hop={args.hop} (BASELINE configs[{(2 if world == 1 else 3) if not geom else 4}])"),
                   "clips_total": total_clips, "clips_per_gpu": n_clips, "clip_seconds": args.clip_seconds, "input": "int16 PCM",
                   "fft": args.fft, "distinct_clips": n_base,
                   "l2": "inputs (%.1f GB per GPU per step) are far larger than the 126 MB L2" % (plan.nS * 2 / 1e9),
                   "collective": "all_gather of per-clip stats (32 B/clip)" if world > 1 else "none (1 GPU)"},
        "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches_per_step * args.steps),
    }

    if not args.no_e2e:
        # ---- end to end through the plugin: RainDetectorProcessor.run_batch on host arrays, default flags
        from audio_processing_tools_b200.edge.rain_signal_processor import RainDetectorProcessor
        from audio_processing_tools_b200.synth import pcm_to_f32
        del pcm_dev
        for k in list(bufs):
            del bufs[k]
        torch.cuda.empty_cache()
        prev_affinity = bind_near_gpu(dev.index) if world > 1 else None   # staging buffers on the GPU's own NUMA node
        # The loader's layout (Mark3BatchLoader, parse.py): the PCM of the batch packed into ONE pinned host buffer, each clip
        # a view of it -- the contract's "inputs from pinned host memory".  The second leg hands the plugin 1 000 separate
        # pageable arrays instead (what an arbitrary caller has), which costs a staging pass through host memory.
        host = torch.empty(plan.nS, dtype=torch.int16, pin_memory=True)
        hv = host.numpy().reshape(n_clips, N)
        for j, i in enumerate(range(lo, hi)):
            hv[j] = base[i % n_base]
        if args.e2e_input == "int16":
            pinned_clips = [hv[j] for j in range(n_clips)]
            pageable_clips = [np.array(base[i % n_base], copy=True) for i in range(lo, hi)]
        else:
            hostf = torch.empty(plan.nS, dtype=torch.float32, pin_memory=True)
            hf = hostf.numpy().reshape(n_clips, N)
            for j, i in enumerate(range(lo, hi)):
                hf[j] = pcm_to_f32(base[i % n_base])
            pinned_clips = [hf[j] for j in range(n_clips)]
            pageable_clips = [np.array(hf[j], copy=True) for j in range(n_clips)]
        esz = pinned_clips[0].itemsize
        proc = RainDetectorProcessor(device=local_rank, fft_f64=FFT_MODE[args.fft])

        def plugin_leg(host_clips):
            outs = proc.run_batch(host_clips, params)          # warm (plan, pinned ring, staging buffers)
            del outs
            if world > 1:
                dist.barrier()
            t_pl, t_call = [], []
            for _ in range(args.e2e_steps):
                t0 = time.perf_counter()
                outs = proc.run_batch(host_clips, params)
                rows_local = torch.tensor(np.asarray([[i, m["rain_frame_count"], m["clip_rain_fraction"], float(m["clip_is_rain"]),
                                                       m["clip_rain_conf"], m["median_rain_conf"], 0.0, 0.0]
                                                      for i, (m, _) in enumerate(outs)], dtype=np.float32))
                if world > 1:
                    gather_clip_stats(rows_local.to(dev), counts, clip_id_base=0)
                    torch.cuda.synchronize()
                t_pl.append(time.perf_counter() - t0)
                t_call.append(proc.last_host_call_s)
                del outs
            dt = float(np.mean(t_pl))
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return dt, float(np.mean(t_call))

        d2h = int(nF * (1 + 4 + 4) + n_clips * (4 + 32))
        for key, clips_, what in (("e2e", pinned_clips, "views of one pinned host buffer (the loader's layout): copied host->device group by group without staging"),
                                  ("e2e_pageable", pageable_clips, "separate pageable host arrays: apt_run_host_clips stages each clip group into a pinned ring "
                                                                   "with helper threads first")):
            dt, dt_call = plugin_leg(clips_)
            result[key] = {"value": total_clips * args.clip_seconds / dt, "unit": "audio-s/s",
                           "h2d_bytes_per_step": int(plan.nS * esz) * world if not strong else int(total_clips * N * esz),
                           "d2h_bytes_per_step": d2h * world if not strong else int(total_clips * (T * 9 + 36)),
                           "ms_per_step": dt * 1e3, "n_gpus": world, "input": args.e2e_input,
                           "ms_in_c_abi_call": dt_call * 1e3, "ms_python_packaging": (dt - dt_call) * 1e3 if world == 1 else None,
                           "note": "RainDetectorProcessor.run_batch(list of %d host arrays, default flags) on every rank; the arrays are %s; H2D / compute / "
                                   "D2H pipelined over clip groups, results land in caller-owned pinned arrays; the timed region includes the "
                                   "per-clip result / state dictionaries; wall clock, max over ranks" % (n_clips, what)
                                   + ("; each rank bound to its GPU's NUMA node" if prev_affinity is not None else "")}
        del proc, pinned_clips, pageable_clips
        # ---- the same through the C ABI alone: the pinned PCM buffer -> apt_run_host_i16 -> pinned result arrays
        outs = {"frame_class": torch.empty(nF, dtype=torch.int8, pin_memory=True).numpy(),
                "event_count": torch.empty(n_clips, dtype=torch.int32, pin_memory=True).numpy(),
                "clip_stats": torch.empty((n_clips, 8), dtype=torch.float32, pin_memory=True).numpy(),
                "rain_conf": None, "noise_conf": None, "event_idx": torch.empty(nF, dtype=torch.int32, pin_memory=True).numpy()}
        eng.run_host_i16(plan, host.numpy(), outs)      # warm (allocates the staging buffers)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            eng.run_host_i16(plan, host.numpy(), outs)
            if world > 1:
                gather_clip_stats(torch.from_numpy(outs["clip_stats"]).to(dev), counts, clip_id_base=0)
                torch.cuda.synchronize()
        dta = (time.perf_counter() - t0) / args.e2e_steps
        if world > 1:
            t = torch.tensor([dta], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dta = float(t.item())
        if prev_affinity is not None:
            os.sched_setaffinity(0, prev_affinity)      # the CPU baseline below uses every host thread
        result["e2e_abi"] = {"value": total_clips * args.clip_seconds / dta, "unit": "audio-s/s", "ms_per_step": dta * 1e3,
                             "h2d_bytes_per_step": int(total_clips * N * 2) if strong else int(plan.nS * 2) * world,
                             "d2h_bytes_per_step": (int(nF * (1 + 4) + n_clips * (4 + 32)) * world),
                             "note": "apt_run_host_i16 on every rank: one pinned host PCM buffer, clip groups pipelined H2D / compute / D2H"}
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        v, n, dt = cpu_port(params, args.clip_seconds, cores)
        result["cpu_baseline"] = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "port",
                                  "sample": f"{n} x {args.clip_seconds:g}s clips, full pipeline, C port of the reference "
                                            f"algorithm on {cores} threads ({dt:.1f}s wall)"}
        if not args.no_cpu_reference and reference_available():
            workers = max(1, cores - 1)
            vr, dtr = cpu_reference(60.0, workers, workers)
            result["cpu_reference"] = {"value": vr, "unit": "audio-s/s", "cores": workers, "kind": "reference",
                                       "sample": f"{workers} x 60s clips, the UNMODIFIED reference (baseline/_ref) through its own "
                                                 f"process_audio_batches_v2 + RainDetectorProcessor, parallel=True, {workers} worker "
                                                 f"processes ({dtr:.1f}s wall); librosa stand-in from oracle/refharness"}
    if rank == 0:
        print(json.dumps(result), flush=True)      # before any teardown: a line lost in a buffer is a lost measurement
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
