import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
from audio_processing_tools_b200.config import build_noise_config
from audio_processing_tools_b200.engine import BatchEngine
from audio_processing_tools_b200.synth import default_params, synth_clip_i16, FS
clips = [synth_clip_i16(s, 700 + i, lam) for i, (s, lam) in enumerate(((60.0, 3.0), (7.3, 3.0), (33.1, 10.0), (20.6, 0.0), (3.2, 10.0)))]
params = default_params(check_duration=3)
eng = BatchEngine(build_noise_config(FS, params), FS)
plan, fast = eng.run_clips(clips, ("gate", "td_fast_crest"))
plan, exact = eng.run_clips(clips, ("gate", "td"))
c64, c32 = exact["td"][0], fast["td_fast_crest"]
bad = np.flatnonzero(fast["gate"] != exact["gate"])
print("gate mismatches", bad.size, "of", c64.size)
for c in range(plan.n_clips):
    f0, f1 = int(plan.frame_off[c]), int(plan.frame_off[c + 1])
    T = f1 - f0
    d = np.abs(c32[f0:f1] - c64[f0:f1]) / np.maximum(c64[f0:f1], 1e-9)
    ok = c64[f0:f1] > 0
    big = np.flatnonzero(ok & (d > 1e-4))
    b = bad[(bad >= f0) & (bad < f1)] - f0
    print(f"clip {c}: T={T} tiles={max(1,(T-2+55)//56)} gate-mismatch frames {b[:12]} (tile {b[:12]//56}); crest dev>1e-4 at {big[:12]} maxdev {d[ok].max():.3e}")
    for t in b[:4]:
        print("   frame", t, "crest32", c32[f0+t], "crest64", c64[f0+t], "gate fast/exact", fast["gate"][f0+t], exact["gate"][f0+t])
