"""TD gate stage (float32 kernel + exact re-check) against the guard band: duration of the TD launches of one step, alone
on the GPU (two-phase schedule, one time segment)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_processing_tools_b200.config import build_noise_config
from audio_processing_tools_b200.engine import BatchEngine
from audio_processing_tools_b200.synth import default_params, synth_clip_i16, batch_clip_spec

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 500
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 600.0
os.environ["APT_SEGMENTS"] = "1"
os.environ["APT_TWO_PHASE"] = "1"
params = default_params(check_duration=seconds)
base = [synth_clip_i16(seconds, *batch_clip_spec(i)) for i in range(8)]
N = base[0].size
pcm = torch.from_numpy(np.stack(base)).cuda().repeat((n_clips + 7) // 8, 1)[:n_clips].contiguous().reshape(-1)
for guard, fast in (("0", "1"), ("1e-4", "1"), ("5e-4", "1"), ("1e-3", "1"), ("1e-3", "0")):
    os.environ["APT_TD_GUARD"] = guard
    os.environ["APT_TD_FAST"] = fast
    eng = BatchEngine(build_noise_config(11162, params), 11162)
    plan = eng.plan_for([N] * n_clips)
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
    print(f"fast={fast} guard={guard:>5}: stft {a[0,0,1]-a[0,0,0]:7.2f} ms  td {a[1,0,1]-a[1,0,0]:7.2f} ms  step {tot.value:7.2f} ms  rain frames {int(bufs['event_count'].sum())}")
    eng.close()
    del bufs, plan, eng
