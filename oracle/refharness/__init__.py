"""Import harness for the UNMODIFIED reference (test infrastructure only).

`install()` puts a librosa stand-in and import-time stubs for the reference's
absent third-party modules on sys.path / sys.modules, then makes
`audio_processing_tools` (the reference package) importable from
$APT_REFERENCE, <repo>/baseline/_ref (the offline `pip install --target` copy of the
unmodified reference, git-ignored; it travels to the GPU box) or /root/reference.  It is
used by oracle/make_golden*.py (build container) and by bench.py's CPU reference legs --
never by the product package.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from unittest import mock

_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_root():
    repo = os.path.dirname(os.path.dirname(_HERE))
    for cand in (os.environ.get("APT_REFERENCE"), os.path.join(repo, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "audio_processing_tools")):
            return cand
    return None


def _stub(name):
    m = mock.MagicMock(name=name)
    m.__name__ = name
    m.__path__ = []
    m.__spec__ = None
    sys.modules[name] = m
    return m


def install():
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not found (set APT_REFERENCE)")
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)          # -> `import librosa` finds the stand-in
    if root not in sys.path:
        sys.path.insert(1, root)
    for name in ("sqlalchemy", "sqlalchemy.dialects", "sqlalchemy.dialects.postgresql",
                 "sqlalchemy.engine", "sqlalchemy.engine.base", "boto3", "botocore",
                 "matplotlib", "matplotlib.pyplot", "matplotlib.widgets", "matplotlib.ticker",
                 "IPython", "IPython.display", "plotly", "plotly.graph_objects",
                 "plotly.subplots", "librosa.display", "tqdm.notebook"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _stub(name)
    if "botocore.exceptions" not in sys.modules:
        be = types.ModuleType("botocore.exceptions")

        class ClientError(Exception):
            pass

        class NoCredentialsError(Exception):
            pass

        be.ClientError = ClientError
        be.NoCredentialsError = NoCredentialsError
        be.__getattr__ = lambda n: type(n, (Exception,), {})
        sys.modules["botocore.exceptions"] = be
    if "kaitaistruct" not in sys.modules:
        ks = types.ModuleType("kaitaistruct")

        class KaitaiStruct:
            def __init__(self, _io=None, _parent=None, _root=None):
                self._io = _io

        class ValidationNotEqualError(Exception):
            pass

        ks.KaitaiStruct = KaitaiStruct
        ks.ValidationNotEqualError = ValidationNotEqualError
        ks.KaitaiStream = mock.MagicMock(name="KaitaiStream")
        ks.BytesIO = __import__("io").BytesIO
        ks.__version__ = "0.10"
        ks.API_VERSION = (0, 10)
        sys.modules["kaitaistruct"] = ks
    import audio_processing_tools  # noqa: F401  (the reference package)
    try:
        emu = importlib.import_module(
            "audio_processing_tools.host_analysis.device_dsd_processing_emulator")
        sys.modules["audio_processing_tools.edge.device_dsd_processing_emulator"] = emu
    except Exception:
        pass
    return root
