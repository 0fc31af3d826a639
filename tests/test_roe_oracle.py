"""Legacy RoE rain detector: the numpy oracle against outputs of the unmodified reference."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from audio_processing_tools_b200.synth import pcm_to_f32, quiet_clip_i16, synth_clip_i16

INT_KEYS = ("rain_drops", "rain_drop_count", "rain_peaks_count", "rain_drop_count_mod", "max_harmonics_out")
ARRAY_KEYS = ("raining", "kurtosis", "crest_factor", "diff_energy", "Nov0")


def roe_cases():
    g = np.load(os.path.join(GOLDEN_DIR, "roe_cases.npz"))
    return g, json.loads(str(g["meta"])), json.loads(str(g["default_params"]))


def case_pcm(m):
    pcm = synth_clip_i16(m["seconds"], m["seed"], m["arg"]) if m["kind"] == "synth" else quiet_clip_i16(m["seconds"], m["seed"], tuple(m["arg"]))
    assert hashlib.sha1(pcm.tobytes()).hexdigest() == m["pcm_sha1"]
    return pcm


def check_roe(out, g, name, rtol):
    """out = (rain_drops, frain_mean, state, max_harmonics_out).  Counts and the per-frame rain status are bit-exact;
    float series within rtol (float64 FFT / summation-order differences)."""
    drops, frain_mean, st, mh = out
    sc = json.loads(str(g[f"{name}__scalars"]))
    got = {"rain_drops": drops, "rain_drop_count": st["rain_drop_count"], "rain_peaks_count": st["rain_peaks_count"],
           "rain_drop_count_mod": st["rain_drop_count_mod"], "max_harmonics_out": mh}
    for k in INT_KEYS:
        assert int(got[k]) == sc[k], k
    assert float(frain_mean) == pytest.approx(sc["frain_mean"], rel=1e-12)
    ref_rain = g[f"{name}__raining"]
    assert np.array_equal(np.asarray(st["raining"]) >= 1, ref_rain >= 1)
    for k in ARRAY_KEYS:
        if f"{name}__{k}" not in g.files:      # no time-domain series when both handle_fp and handle_fn are off
            assert k not in st
            continue
        np.testing.assert_allclose(np.asarray(st[k], np.float64), g[f"{name}__{k}"], rtol=rtol, atol=1e-12, err_msg=k)


@pytest.mark.parametrize("idx", range(10))
def test_roe_oracle_matches_reference(idx):
    from oracle import roe_oracle
    g, meta, defaults = roe_cases()
    m = meta[idx]
    sc = json.loads(str(g[m["name"] + "__scalars"]))
    params = dict(defaults)
    params.update(m["extra"])
    out = roe_oracle.rain_detection_algo(pcm_to_f32(case_pcm(m)), max_harmonics=sc["max_harmonics_in"], **params)
    check_roe(out, g, m["name"], rtol=1e-9)
