"""Runs the device-resident step back to back for a few seconds (for an nvidia-smi clock / power log beside it)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_processing_tools_b200.config import build_noise_config
from audio_processing_tools_b200.engine import BatchEngine
from audio_processing_tools_b200.synth import default_params, synth_clip_i16, batch_clip_spec
n_clips, seconds, secs = int(sys.argv[1]), float(sys.argv[2]), float(sys.argv[3])
params = default_params(check_duration=seconds)
eng = BatchEngine(build_noise_config(11162, params), 11162)
base = [synth_clip_i16(seconds, *batch_clip_spec(i)) for i in range(8)]
N = base[0].size
plan = eng.plan_for([N] * n_clips)
pcm = torch.from_numpy(np.stack(base)).cuda().repeat((n_clips + 7) // 8, 1)[:n_clips].contiguous().reshape(-1)
bufs = eng.alloc_outputs(plan, (), full=True)
torch.cuda.synchronize()
print("start", time.time(), flush=True)
t0 = time.time(); n = 0
while time.time() - t0 < secs:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.run_device(plan, pcm, bufs, full=True); e1.record(); torch.cuda.synchronize()
    if n % 5 == 0: print(f"t={time.time()-t0:5.2f}s step {e0.elapsed_time(e1):.2f} ms", flush=True)
    n += 1
print("end", time.time())
