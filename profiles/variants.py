"""Schedules of the pipelined run, device-resident batch: step time for each knob combination (experiments)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_processing_tools_b200.config import build_noise_config
from audio_processing_tools_b200.engine import BatchEngine
from audio_processing_tools_b200.synth import default_params, synth_clip_i16, batch_clip_spec

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 600.0
params = default_params(check_duration=seconds)
eng = BatchEngine(build_noise_config(11162, params), 11162)
base = [synth_clip_i16(seconds, *batch_clip_spec(i)) for i in range(8)]
N = base[0].size
plan = eng.plan_for([N] * n_clips)
pcm = torch.from_numpy(np.stack(base)).cuda().repeat((n_clips + 7) // 8, 1)[:n_clips].contiguous().reshape(-1)
bufs = eng.alloc_outputs(plan, (), full=True)
ref = None
for name, env in (("pipelined 16 seg", {}),
                  ("td own stream", {"APT_TD_OWN_STREAM": "1"}),
                  ("two-phase", {"APT_TWO_PHASE": "1"}),
                  ("three-phase", {"APT_TWO_PHASE": "2"}),
                  ("three-phase 8 seg", {"APT_TWO_PHASE": "2", "APT_SEGMENTS": "8"}),
                  ("three-phase 32 seg", {"APT_TWO_PHASE": "2", "APT_SEGMENTS": "32"}),
                  ("two-phase + td own stream", {"APT_TWO_PHASE": "1", "APT_TD_OWN_STREAM": "1"}),
                  ("two-phase 8 seg", {"APT_TWO_PHASE": "1", "APT_SEGMENTS": "8"}),
                  ("two-phase 32 seg", {"APT_TWO_PHASE": "1", "APT_SEGMENTS": "32"}),
                  ("pipelined 4 seg", {"APT_SEGMENTS": "4"}),
                  ("one segment (serial)", {"APT_SEGMENTS": "1"}),
                  ("serial, exact TD", {"APT_SEGMENTS": "1", "APT_TD_FAST_OFF": "1"})):
    for k in ("APT_TD_OWN_STREAM", "APT_TWO_PHASE", "APT_SEGMENTS"):
        os.environ.pop(k, None)
    os.environ.update({k: v for k, v in env.items() if k != "APT_TD_FAST_OFF"})
    for _ in range(2):
        eng.run_device(plan, pcm, bufs, full=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        eng.run_device(plan, pcm, bufs, full=True)
    e1.record()
    torch.cuda.synchronize()
    fc = bufs["frame_class"].clone()
    if ref is None:
        ref = fc
    print(f"{name:28s} {e0.elapsed_time(e1) / 3:8.2f} ms/step   labels equal: {bool((fc == ref).all())}")
