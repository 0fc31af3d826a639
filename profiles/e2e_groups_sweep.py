"""End-to-end time of apt_run_host_i16 as a function of the clip-group count (APT_HOST_GROUPS); prints one line per setting."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_processing_tools_b200.config import build_noise_config
from audio_processing_tools_b200.engine import BatchEngine
from audio_processing_tools_b200.synth import default_params, synth_clip_i16, batch_clip_spec
n_clips, seconds = 1000, 600.0
params = default_params(check_duration=seconds)
eng = BatchEngine(build_noise_config(11162, params), 11162)
base = [synth_clip_i16(seconds, *batch_clip_spec(i)) for i in range(4)]
N = base[0].size
plan = eng.plan_for([N] * n_clips)
host = torch.empty(plan.nS, dtype=torch.int16, pin_memory=True)
hv = host.numpy().reshape(n_clips, N)
for i in range(n_clips): hv[i] = base[i % 4]
dev = torch.empty(plan.nS, dtype=torch.int16, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); dev.copy_(host, non_blocking=True); torch.cuda.synchronize()
    print("raw H2D 13.4 GB: %.1f ms -> %.1f GB/s" % ((time.perf_counter() - t0) * 1e3, plan.nS * 2 / (time.perf_counter() - t0) / 1e9), flush=True)
del dev
nF = plan.nF
outs = {"frame_class": torch.empty(nF, dtype=torch.int8, pin_memory=True).numpy(),
        "event_count": torch.empty(n_clips, dtype=torch.int32, pin_memory=True).numpy(),
        "clip_stats": torch.empty((n_clips, 8), dtype=torch.float32, pin_memory=True).numpy(),
        "rain_conf": None, "noise_conf": None, "event_idx": torch.empty(nF, dtype=torch.int32, pin_memory=True).numpy()}
for g in (4, 8, 12, 16, 24, 32):
    os.environ["APT_HOST_GROUPS"] = str(g)
    eng.run_host_i16(plan, host.numpy(), outs)
    t0 = time.perf_counter()
    for _ in range(2): eng.run_host_i16(plan, host.numpy(), outs)
    print("groups", g, "e2e ms", (time.perf_counter() - t0) / 2 * 1e3, flush=True)
