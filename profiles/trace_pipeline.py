"""Timeline of the pipelined run: per kernel kind and time segment, start / end (ms since the start of the step) from
timing events on the kind's stream (apt_plan_trace).  usage: python profiles/trace_pipeline.py [clips] [seconds]"""
import ctypes as C, os, sys
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
for _ in range(2):
    eng.run_device(plan, pcm, bufs, full=True)
torch.cuda.synchronize()
eng.L.apt_plan_enable_trace(plan.h, 1)
eng.run_device(plan, pcm, bufs, full=True)
torch.cuda.synchronize()
out = (C.c_float * (8 * 64 * 2))()
nseg, tot = C.c_int(0), C.c_float(0)
eng.L.apt_plan_trace(plan.h, out, C.byref(nseg), C.byref(tot))
a = np.frombuffer(out, dtype=np.float32).reshape(8, 64, 2)
kinds = ("stft", "td", "trk1", "flux", "base", "decide", "trk2", "dbsum")
print(f"clips={n_clips} seconds={seconds} segments={nseg.value} total_ms={tot.value:.3f}")
print("kind    " + " ".join(f"{'s%d' % s:>13}" for s in range(nseg.value)))
for k, name in enumerate(kinds):
    print(f"{name:8}" + " ".join(f"{a[k, s, 0]:6.2f}-{a[k, s, 1]:6.2f}" for s in range(nseg.value)))
print("busy ms per kind: " + ", ".join(f"{name} {float((a[k, :nseg.value, 1] - a[k, :nseg.value, 0]).sum()):.2f}" for k, name in enumerate(kinds)))
