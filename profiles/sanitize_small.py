"""Tiny end-to-end exercise of every kernel for compute-sanitizer memcheck."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from audio_processing_tools_b200.config import build_noise_config
from audio_processing_tools_b200.engine import BatchEngine
from audio_processing_tools_b200.synth import default_params, synth_clip_i16, quiet_clip_i16
from audio_processing_tools_b200.host_analysis.device_dsd_processing_emulator import DsdProcessingEmualtor
ALL = ("S", "P", "det_noise_psd", "det_noise_lag", "D", "noise_psd", "mode_flux", "norm_flux", "score", "td", "raw",
       "band_energy", "gate", "x_td", "G", "ratio_med", "S_hat")
# ragged: 45 s and 25.3 s clips span several time segments of the pipelined run (1792 frames = 20.5 s each), 3.7 s does not
clips = [synth_clip_i16(sec, 900 + i, 10.0) for i, sec in enumerate((45.0, 25.3, 3.7))]
params = default_params(check_duration=3)
eng = BatchEngine(build_noise_config(11162, params), 11162)
plan, out = eng.run_clips(clips, ALL)
plan, out2 = eng.run_clips(clips, ())
plan, out3 = eng.run_clips([c.astype(np.float32) / np.float32(32767) for c in clips], ())
assert np.array_equal(out["frame_class"], out2["frame_class"]) and np.array_equal(out2["frame_class"], out3["frame_class"])
host = {"frame_class": np.zeros(plan.nF, np.int8), "rain_conf": np.zeros(plan.nF, np.float32), "noise_conf": np.zeros(plan.nF, np.float32),
        "event_idx": np.zeros(plan.nF, np.int32), "event_count": np.zeros(3, np.int32), "clip_stats": np.zeros((3, 8), np.float32)}
eng.run_host_i16(plan, np.concatenate(clips), host)
assert np.array_equal(host["frame_class"], out2["frame_class"])
plan4, out4 = eng.run_host_clips([c.copy() for c in clips], event_idx=True)
assert np.array_equal(out4["frame_class"], out2["frame_class"]) and np.array_equal(out4["clip_stats"], out2["clip_stats"])
plan5, out5 = eng.run_host_clips([c.astype(np.float32) / np.float32(32767) for c in clips])
assert np.array_equal(out5["frame_class"], out2["frame_class"])
eng.close()
for n_fft, hop in ((256, 64), (512, 128), (4096, 1024)):
    p2 = default_params(check_duration=3, n_fft=n_fft, hop=hop)
    e2 = BatchEngine(build_noise_config(11162, p2), 11162)
    e2.run_clips(clips, ("S", "P", "raw", "band_energy"), full=False)
    e2.close()
DsdProcessingEmualtor().process_audio_batch([synth_clip_i16(70, 1, 3.0), quiet_clip_i16(65, 2, (30.0,))], [0, 17])
print("sanitize-small ok")
