#!/usr/bin/env python
"""Golden vectors of the FULL detector at frame sizes other than 256 / 128 (BASELINE config 5 with the whole pipeline).

Runs the UNMODIFIED reference (RainDetectorProcessor.run with n_fft / hop overrides, keep_state_debug) in the build
container and freezes labels, confidences, clip statistics and the per-frame detector features.  The PCM is regenerated
from the seed by the tests (sha1 recorded).  Test infrastructure; see make_golden.py for the harness.

    python oracle/make_golden_geom.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)

import make_golden as mg  # noqa: E402  (installs the reference harness)
from audio_processing_tools_b200.synth import synth_clip_i16  # noqa: E402

# (n_fft, hop, seconds, seed, rain rate): every hop is a multiple of 64 (the CUDA path's condition for the full pipeline)
CASES = ((512, 256, 30, 51, 3.0), (1024, 256, 30, 52, 10.0), (2048, 1024, 40, 53, 3.0), (4096, 1024, 40, 54, 3.0),
         (256, 256, 20, 55, 3.0), (512, 128, 20, 56, 0.5),
         # hops that are multiples of 64 only (13 s: N mod 128 = 82, a 128-sample leaf fits behind the last whole block)
         (256, 64, 13, 57, 3.0), (512, 192, 13, 58, 10.0))


def main():
    for n_fft, hop, seconds, seed, lam in CASES:
        pcm = synth_clip_i16(seconds, seed, lam)
        metrics, state, params = mg.run_reference(pcm, seconds, {"n_fft": n_fft, "hop": hop})
        d = mg.pack(pcm, seconds, seed, lam, metrics, state, level=1, params=params)
        path = os.path.join(mg.OUT, f"geom_nfft{n_fft}_hop{hop}.npz")
        np.savez_compressed(path, **d)
        fc = d["frame_class"]
        print(path, fc.size, "frames,", int((fc == 2).sum()), "rain,", os.path.getsize(path) // 1024, "KiB", flush=True)


def features_case():
    """`dump_features` payload (state["features"]) at feature_decim = 3, default detector flags."""
    import hashlib
    import json
    from audio_processing_tools.edge.rain_signal_processor import RainDetectorProcessor
    from audio_processing_tools_b200.synth import default_params, pcm_to_f32
    seconds, seed, lam = 8, 61, 3.0
    pcm = synth_clip_i16(seconds, seed, lam)
    params = default_params(check_duration=seconds, dump_features=True, feature_decim=3)
    _, state = RainDetectorProcessor().run(pcm_to_f32(pcm), params)
    d = {"feat_" + k: np.asarray(v) for k, v in state["features"].items()}
    d["meta"] = np.array(json.dumps({"seconds": seconds, "seed": seed, "lam": lam, "feature_decim": 3,
                                     "pcm_sha1": hashlib.sha1(pcm.tobytes()).hexdigest(), "numpy": np.__version__}))
    path = os.path.join(mg.OUT, "features_s61_decim3.npz")
    np.savez_compressed(path, **d)
    print(path, sorted(d), flush=True)


def adaptive_q_cases():
    """adaptive_q_enable (rain_signal_processor.py:570-576, :634-638): level-2 fixtures (every exported array), appended
    to INDEX.json so that the oracle and GPU parity tests pick them up with the other fixtures."""
    import json
    idx_path = os.path.join(mg.OUT, "INDEX.json")
    with open(idx_path) as f:
        index = json.load(f)
    for name, seconds, seed, lam, extra in (
            ("alt_s33_l3_10s_adaptq", 10, 33, 3.0, {"adaptive_q_enable": True}),
            ("alt_s34_l10_10s_adaptq", 10, 34, 10.0, {"adaptive_q_enable": True, "adaptive_q_min": 0.05, "adaptive_q_alpha": 0.9, "q": 0.3}),
            # spectral SNR gating of the oversubtraction (:1050-1077): seen in the gain plane G_band
            ("alt_s39_l3_10s_snrgate", 10, 39, 3.0, {"snr_gating_enable": True, "snr_gating_snr1": 2.0}),
            # the same with the power step on the gate (:1073-1075) over the whole operating band
            ("alt_s40_l10_10s_snrgate_pow", 10, 40, 10.0, {"snr_gating_enable": True, "snr_gating_snr1": 0.5, "snr_gating_power": 2.5,
                                                           "snr_gating_use_mode_bands": False})):
        pcm = synth_clip_i16(seconds, seed, lam)
        metrics, state, params = mg.run_reference(pcm, seconds, extra, spectra=True)
        d = mg.pack(pcm, seconds, seed, lam, metrics, state, level=2, params=params)
        if state["debug"].get("snr_gate") is not None:
            d["snr_gate_t"] = np.asarray(state["debug"]["snr_gate"])
            d["snr_mode_t"] = np.asarray(state["debug"]["snr_mode"])
        np.savez_compressed(os.path.join(mg.OUT, name + ".npz"), **d)
        fc = d["frame_class"]
        entry = {"name": name, "seconds": seconds, "seed": seed, "lam": lam, "level": 2, "T": int(fc.size),
                 "rain": int((fc == 2).sum()), "uncertain": int((fc == 1).sum()), "noise": int((fc == 0).sum()), "extra": extra}
        index = [e for e in index if e["name"] != name] + [entry]
        print(entry, os.path.getsize(os.path.join(mg.OUT, name + ".npz")) // 1024, "KiB", flush=True)
    with open(idx_path, "w") as f:
        json.dump(index, f, indent=1)


def smoothing_cases():
    """pre_smooth_frames / median_frames (rain_signal_processor.py:366-396, :690-692, :717-719): level-1 fixtures (labels,
    per-frame detector features, clip statistics); both options move the labels through the detector's noise baseline."""
    import json
    idx_path = os.path.join(mg.OUT, "INDEX.json")
    with open(idx_path) as f:
        index = json.load(f)
    for name, seconds, seed, lam, extra in (
            ("alt_s35_l3_20s_presmooth3", 20, 35, 3.0, {"pre_smooth_frames": 3}),
            ("alt_s36_l10_20s_median5", 20, 36, 10.0, {"median_frames": 5}),
            ("alt_s37_l3_20s_smooth_all", 20, 37, 3.0, {"pre_smooth_frames": 4, "median_frames": 4, "adaptive_q_enable": True}),
            # bypass_classifier (:846-857): every frame NOISE, the suppressor's noise PSD alone (level 0: labels + clip statistics)
            ("alt_s40_l10_20s_bypasscls", 20, 40, 10.0, {"detector": {"bypass_classifier": True}})):
        pcm = synth_clip_i16(seconds, seed, lam)
        metrics, state, params = mg.run_reference(pcm, seconds, extra)
        level = 0 if "detector" in extra else 1
        d = mg.pack(pcm, seconds, seed, lam, metrics, state, level=level, params=params)
        np.savez_compressed(os.path.join(mg.OUT, name + ".npz"), **d)
        fc = d["frame_class"]
        entry = {"name": name, "seconds": seconds, "seed": seed, "lam": lam, "level": level, "T": int(fc.size),
                 "rain": int((fc == 2).sum()), "uncertain": int((fc == 1).sum()), "noise": int((fc == 0).sum()), "extra": extra}
        index = [e for e in index if e["name"] != name] + [entry]
        print(entry, os.path.getsize(os.path.join(mg.OUT, name + ".npz")) // 1024, "KiB", flush=True)
    with open(idx_path, "w") as f:
        json.dump(index, f, indent=1)


def feature_dump_case():
    """dump_features with the detector-side dump on (feature_dump_level = 1, dense + sparse, soft TD votes), decimated by 2:
    every array of state["features"] (rain_frame_classifier.py:1096-1162, rain_signal_processor.py:723-787)."""
    import hashlib
    import json
    from audio_processing_tools.edge.rain_signal_processor import RainDetectorProcessor
    from audio_processing_tools_b200.synth import default_params, pcm_to_f32
    seconds, seed, lam = 8, 64, 10.0
    pcm = synth_clip_i16(seconds, seed, lam)
    det = {"feature_dump_level": 1, "feature_dump_sparse_enable": True, "feature_dump_sparse_gate_threshold": 3.0,
           "feature_dump_include_td_soft": True, "td_soft_enable": True}
    params = default_params(check_duration=seconds, dump_features=True, feature_decim=2)
    params["detector"] = dict(params["detector"], **det)
    _, state = RainDetectorProcessor().run(pcm_to_f32(pcm), params)
    d = {"feat_" + k: np.asarray(v) for k, v in state["features"].items()}
    d["meta"] = np.array(json.dumps({"seconds": seconds, "seed": seed, "lam": lam, "feature_decim": 2, "detector": det,
                                     "pcm_sha1": hashlib.sha1(pcm.tobytes()).hexdigest(), "numpy": np.__version__}))
    path = os.path.join(mg.OUT, "featuredump_s64.npz")
    np.savez_compressed(path, **d)
    print(path, len(d) - 1, "arrays", {k: v.shape for k, v in d.items() if k != "meta"}, flush=True)


def detdebug_case():
    """The full key set of the detector's debug dictionary (state["det_debug"]) with the scalar echoes, and the soft TD
    label arrays under td_soft_enable (rain_frame_classifier.py:85-110, :1000-1047)."""
    import hashlib
    import json
    from audio_processing_tools.edge.rain_signal_processor import RainDetectorProcessor
    from audio_processing_tools_b200.synth import default_params, pcm_to_f32
    seconds, seed, lam = 8, 63, 10.0
    pcm = synth_clip_i16(seconds, seed, lam)
    params = default_params(check_duration=seconds, keep_state_debug=True)
    params["detector"] = dict(params["detector"], td_soft_enable=True, td_soft_crest_factor_min=3.0, td_soft_kurtosis_min=4.0)
    _, state = RainDetectorProcessor().run(pcm_to_f32(pcm), params)
    dd = state["det_debug"]
    scalars = {k: (v if not isinstance(v, (np.floating, np.integer, np.bool_)) else v.item())
               for k, v in dd.items() if not isinstance(v, np.ndarray) and not isinstance(v, dict)}
    d = {"meta": np.array(json.dumps({"seconds": seconds, "seed": seed, "lam": lam,
                                      "pcm_sha1": hashlib.sha1(pcm.tobytes()).hexdigest(), "numpy": np.__version__,
                                      "keys": sorted(dd.keys()), "scalars": scalars,
                                      "debug_keys": sorted(state["debug"].keys()), "state_keys": sorted(state.keys())}, default=str)),
         "td_vote_count": np.asarray(dd["td_vote_count"]), "td_soft_score": np.asarray(dd["td_soft_score"]),
         "td_soft_label": np.asarray(dd["td_soft_label"]), "raw_spectral_dump_mask": np.asarray(dd["raw_spectral_dump_mask"]),
         "sparse_frame_idx": np.asarray(dd["sparse_frame_idx"])}
    path = os.path.join(mg.OUT, "detdebug_s63.npz")
    np.savez_compressed(path, **d)
    print(path, len(dd), "keys", scalars, flush=True)


if __name__ == "__main__":
    detdebug_case()
    feature_dump_case()
    main()
    features_case()
    adaptive_q_cases()
    smoothing_cases()
