"""Band noise estimator: the numpy oracle against outputs of the unmodified reference."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from audio_processing_tools_b200.synth import pcm_to_f32, quiet_clip_i16, synth_clip_i16

FLOAT_KEYS = ("M_band", "E_band", "N_E", "N_E_raw", "subE", "N_sub", "G_mag", "M_clean", "noise_effective_q",
              "M_band_fft", "E_band_fft", "E_hpf", "times_s")
BOOL_KEYS = ("rain_submask", "fft_rain_frame")
INT_STATS = ("noise_frame_count", "rain_frame_count", "total_frame_count", "noise_buffer_valid_count",
             "noise_buffer_min_valid_count", "noise_buffer_underflow_frame_count", "frames_since_noise_update",
             "noise_learned_subframe_count", "noise_replenish_count")


def band_cases():
    g = np.load(os.path.join(GOLDEN_DIR, "band_noise_cases.npz"))
    return g, json.loads(str(g["meta"]))


def case_pcm(m):
    pcm = synth_clip_i16(m["seconds"], m["seed"], m["arg"]) if m["kind"] == "synth" else quiet_clip_i16(m["seconds"], m["seed"], tuple(m["arg"]))
    assert hashlib.sha1(pcm.tobytes()).hexdigest() == m["pcm_sha1"]
    return pcm


def check_against_golden(out, g, name, rtol):
    for k in BOOL_KEYS:
        assert np.array_equal(np.asarray(out[k], bool), g[f"{name}__{k}"]), k      # decisions: bit-exact
    for k in FLOAT_KEYS:
        np.testing.assert_allclose(np.asarray(out[k], np.float64), g[f"{name}__{k}"], rtol=rtol, atol=1e-300, err_msg=k)
    es = json.loads(str(g[f"{name}__energy_stats"]))
    for k in INT_STATS:
        assert int(out["energy_stats"][k]) == int(es[k]), k
    for k in ("noise_energy_sum", "rain_energy_sum", "total_energy_sum", "noise_effective_q"):
        assert float(out["energy_stats"][k]) == pytest.approx(float(es[k]), rel=max(rtol, 1e-12)), k


@pytest.mark.parametrize("idx", range(6))
def test_band_noise_oracle_matches_reference(idx):
    from oracle import band_noise_oracle
    g, meta = band_cases()
    m = meta[idx]
    params = {"sample_rate": 11162, "check_duration": m["seconds"], **m["extra"]}
    out = band_noise_oracle.run(pcm_to_f32(case_pcm(m)), params)
    check_against_golden(out, g, m["name"], rtol=1e-12)
