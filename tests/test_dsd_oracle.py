"""Drop-size-distribution emulator: the numpy oracle against outputs of the unmodified reference."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from audio_processing_tools_b200.synth import quiet_clip_i16, synth_clip_i16


def dsd_cases():
    g = np.load(os.path.join(GOLDEN_DIR, "dsd_cases.npz"))
    return [(m, g[m["name"] + "__out"]) for m in json.loads(str(g["meta"]))]


def case_pcm(m):
    arg = m["arg"]
    pcm = synth_clip_i16(m["seconds"], m["seed"], arg) if m["kind"] == "synth" else quiet_clip_i16(m["seconds"], m["seed"], tuple(arg))
    assert hashlib.sha1(pcm.tobytes()).hexdigest() == m["pcm_sha1"]
    return pcm


@pytest.mark.parametrize("m,ref", dsd_cases(), ids=[m["name"] for m, _ in dsd_cases()])
def test_dsd_oracle_matches_reference(m, ref):
    from oracle import dsd_oracle
    pcm = case_pcm(m)
    out = dsd_oracle.process_audio_data(pcm.astype(np.float64) / 32768.0, m["ts"], window=m["window"])
    assert len(out) == ref.shape[0]
    assert np.array_equal(np.asarray(out).reshape(ref.shape), ref)     # integer histograms: bit-exact
