"""Small representative workload for ncu captures: 444 clips x 20 s (3 resident CTAs/SM of the
sequential kernel), full pipeline, 1 warm-up pass + 1 measured pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_processing_tools_b200.config import build_noise_config
from audio_processing_tools_b200.engine import BatchEngine
from audio_processing_tools_b200.synth import default_params, synth_clip_i16, batch_clip_spec

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 444
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
fft = sys.argv[3] if len(sys.argv) > 3 else "f64"
params = default_params(check_duration=seconds)
eng = BatchEngine(build_noise_config(11162, params), 11162, fft_f64={"f64": True, "f32": False, "tc": "tc"}[fft])
base = [synth_clip_i16(seconds, *batch_clip_spec(i)) for i in range(8)]
N = base[0].size
plan = eng.plan_for([N] * n_clips)
pcm = torch.from_numpy(np.stack(base)).cuda().repeat((n_clips + 7) // 8, 1)[:n_clips].contiguous().reshape(-1)
bufs = eng.alloc_outputs(plan, (), full=True)
for _ in range(2):
    eng.run_device(plan, pcm, bufs, full=True)
    torch.cuda.synchronize()
print("ok", int(bufs["event_count"].sum()))
