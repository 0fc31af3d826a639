"""Throughput of the engines beside the main path: band noise estimator, legacy RoE detector, DSD emulator.
Wall clock around the public Python call (host buffers in, numpy results out: H2D, kernels, D2H, packaging), after one
warm-up call; synthetic clips.  Prints one JSON line per engine."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_processing_tools_b200.synth import synth_clip_i16, batch_clip_spec
from audio_processing_tools_b200.edge.band_noise_processor import BandNoiseEstimatorProcessor
from audio_processing_tools_b200.edge import dsp_rain_detection as roe

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); t.append(time.perf_counter() - t0)
    return min(t)

n_bne, s_bne = 256, 60.0
clips = [synth_clip_i16(s_bne, *batch_clip_spec(i)) for i in range(8)]
batch = [clips[i % 8] for i in range(n_bne)]
proc = BandNoiseEstimatorProcessor()
dt = timed(lambda: proc.run_batch(batch, {"sample_rate": 11162}))
print(json.dumps({"engine": "band_noise_estimator", "clips": n_bne, "clip_seconds": s_bne, "wall_s": dt, "audio_s_per_s": n_bne * s_bne / dt}))

n_roe, s_roe = 1000, 10.0
clips = [synth_clip_i16(s_roe, *batch_clip_spec(i)) for i in range(8)]
batch = [clips[i % 8] for i in range(n_roe)]
dt = timed(lambda: roe.rain_detection_algo_batch(batch, **roe.default_params))
print(json.dumps({"engine": "legacy_roe", "clips": n_roe, "clip_seconds": s_roe, "wall_s": dt, "audio_s_per_s": n_roe * s_roe / dt}))
