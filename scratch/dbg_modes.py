import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from audio_processing_tools_b200.config import build_noise_config
from audio_processing_tools_b200.engine import BatchEngine
from audio_processing_tools_b200.synth import default_params, synth_clip_i16, pcm_to_f32
from oracle import oracle
oracle.build()
clips = [synth_clip_i16(20, 300 + i, (0.0, 3.0, 10.0)[i]) for i in range(3)]
params = default_params(check_duration=20)
for fft64 in (True, False):
    eng = BatchEngine(build_noise_config(11162, params), 11162, fft_f64=fft64)
    for want in (("mode_flux", "norm_flux", "score"), ("mode_flux", "norm_flux", "score", "D")):
        plan, out = eng.run_clips(clips, want)
        for c, pcm in enumerate(clips):
            m, s = oracle.run(pcm_to_f32(pcm), dict(params, keep_state_debug=True))
            f0, f1 = int(plan.frame_off[c]), int(plan.frame_off[c + 1])
            mf = np.abs(out["mode_flux"][:, f0:f1] - s["mode_flux"]).max()
            nf = np.abs(out["norm_flux"][:, f0:f1] - s["norm_flux"]).max()
            fl = int((out["frame_class"][f0:f1] != s["frame_class"]).sum())
            print("fft64", fft64, "want", len(want), "clip", c, "mode_flux err", mf, "norm err", nf, "flips", fl, flush=True)
    eng.close()
