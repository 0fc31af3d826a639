import json
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def golden_names(min_level=0, max_seconds=1e9):
    with open(os.path.join(GOLDEN_DIR, "INDEX.json")) as f:
        idx = json.load(f)
    return [e["name"] for e in idx if e["level"] >= min_level and e["seconds"] <= max_seconds]


def load_golden(name):
    """Returns (golden npz dict, meta dict, int16 pcm, params) for a committed fixture.

    PCM is regenerated from the normative synthetic generator and checked against the
    sha1 recorded when the reference produced the fixture."""
    import hashlib

    from audio_processing_tools_b200.synth import synth_clip_i16

    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    meta = json.loads(str(g["meta"]))
    pcm = g["pcm"] if "pcm" in g else synth_clip_i16(meta["seconds"], meta["seed"], meta["lam"])
    assert hashlib.sha1(pcm.tobytes()).hexdigest() == meta["pcm_sha1"], \
        "synthetic generator no longer reproduces the fixture's PCM (numpy RNG drift?)"
    params = dict(meta["params"])
    params["detector"] = dict(params["detector"])
    params["detector"]["mode_bands"] = [tuple(b) for b in params["detector"]["mode_bands"]]
    if "operating_band" in params:
        params["operating_band"] = tuple(params["operating_band"])
    return g, meta, pcm, params


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle
