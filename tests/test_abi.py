"""The C-ABI library loads on a CPU-only box and exports every symbol include/apt_b200.h declares.
No compute is called here (there is no GPU); the product path must fail loudly, never fall back."""
import ctypes as C
import os
import re

import pytest

from conftest import REPO
from audio_processing_tools_b200 import _lib


def _declared_symbols():
    hdr = open(os.path.join(REPO, "include", "apt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(apt_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported():
    _lib.build()
    L = _lib.load()
    declared = _declared_symbols()
    assert declared, "no declarations parsed from include/apt_b200.h"
    for sym in declared:
        assert hasattr(L, sym), f"{sym} declared in apt_b200.h but not exported by libapt_b200.so"
    assert sorted(_lib.EXPORTS) == declared, "python binding list and header disagree"


def test_struct_layouts_and_defaults():
    L = _lib.load()
    assert L.apt_abi_version() == _lib.ABI_VERSION
    assert L.apt_sizeof_params() == C.sizeof(_lib.AptParams)
    assert L.apt_sizeof_out() == C.sizeof(_lib.AptOut)
    p = _lib.AptParams()
    assert L.apt_params_default(C.byref(p)) == 0
    # defaults of the reference at fs=11162
    assert (p.fs, p.n_fft, p.hop, p.band_lo, p.band_hi) == (11162, 256, 128, 10, 80)
    assert p.warmup_need == 21 and p.min_support == 2
    assert p.trk_q == pytest.approx(0.25) and p.bl_q == pytest.approx(0.2)
    assert p.thr_primary == pytest.approx(1.8) and p.thr_m3 == pytest.approx(3.0)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is visible")
    L = _lib.load()
    ctx = C.c_void_p()
    assert L.apt_init(0, C.byref(ctx)) < 0 and not ctx.value
    from audio_processing_tools_b200.config import build_noise_config
    from audio_processing_tools_b200.engine import AptError, BatchEngine
    from audio_processing_tools_b200.synth import default_params
    params = default_params()
    with pytest.raises(AptError):
        BatchEngine(build_noise_config(11162, params), 11162)
    from audio_processing_tools_b200.edge.rain_signal_processor import RainDetectorProcessor
    import numpy as np
    with pytest.raises(AptError):
        RainDetectorProcessor().run(np.zeros(11162 * 60, np.float32), params)


def test_product_does_not_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(REPO, "audio_processing_tools_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "apt_oracle" not in src, f
