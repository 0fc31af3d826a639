"""Plugin host path (apt_run_host_clips) against staging threads / clip groups: 1 000 x 600 s int16 clips in pageable memory."""
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
clips = [np.array(base[i % 8], copy=True) for i in range(n_clips)]
print("cpus", os.cpu_count())
for thr, groups in ((8, 24), (12, 24), (16, 24), (24, 24), (32, 24), (16, 12), (16, 48), (16, 96)):
    os.environ["APT_STAGE_THREADS"] = str(thr)
    os.environ["APT_HOST_GROUPS"] = str(groups)
    eng.run_host_clips(clips)
    t0 = time.perf_counter()
    for _ in range(2):
        eng.run_host_clips(clips)
    dt = (time.perf_counter() - t0) / 2
    print(f"threads={thr:3d} groups={groups:3d}  {dt * 1e3:8.1f} ms  {n_clips * seconds / dt / 1e6:6.3f} M audio-s/s")
# plain memcpy bandwidth of this host, one thread and all threads (numpy copies release the GIL)
src = np.concatenate(clips[:64]); dst = np.empty_like(src)
t0 = time.perf_counter(); np.copyto(dst, src); dt = time.perf_counter() - t0
print(f"single-thread numpy copy: {src.nbytes / dt / 1e9:.1f} GB/s")
pin = torch.empty(src.size, dtype=torch.int16, pin_memory=True).numpy()
t0 = time.perf_counter(); np.copyto(pin, src); dt = time.perf_counter() - t0
print(f"single-thread copy into pinned memory: {src.nbytes / dt / 1e9:.1f} GB/s")
