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
        # Stand-in for the un-vendored `kaitaistruct` runtime (uv.lock pins kaitaistruct 0.11), restating the few calls the
        # reference's generated class AudioBinary makes (parse.py:29-54): KaitaiStruct.from_bytes, and on the stream
        # read_bytes / read_u1 / read_u4le / read_f4le / read_bytes_full.  Little-endian fixed-width reads of a byte buffer.
        import struct as _struct
        ks = types.ModuleType("kaitaistruct")

        class KaitaiStream:
            def __init__(self, io_):
                self._io = io_

            def read_bytes(self, n):
                b = self._io.read(n)
                if len(b) < n:
                    raise EOFError(f"requested {n} bytes, but only {len(b)} bytes available")
                return b

            def read_bytes_full(self):
                return self._io.read()

            def read_u1(self):
                return _struct.unpack("<B", self.read_bytes(1))[0]

            def read_u4le(self):
                return _struct.unpack("<I", self.read_bytes(4))[0]

            def read_f4le(self):
                return _struct.unpack("<f", self.read_bytes(4))[0]

        class KaitaiStruct:
            def __init__(self, _io=None, _parent=None, _root=None):
                self._io = _io

            @classmethod
            def from_bytes(cls, buf):
                return cls(KaitaiStream(__import__("io").BytesIO(buf)))

        class ValidationNotEqualError(Exception):
            def __init__(self, expected=None, actual=None, io_=None, src_path=None):
                super().__init__(f"not equal, expected {expected!r}, but got {actual!r} ({src_path})")

        ks.KaitaiStruct = KaitaiStruct
        ks.KaitaiStream = KaitaiStream
        ks.ValidationNotEqualError = ValidationNotEqualError
        ks.BytesIO = __import__("io").BytesIO
        ks.__version__ = "0.11"
        ks.API_VERSION = (0, 11)
        sys.modules["kaitaistruct"] = ks
    import audio_processing_tools  # noqa: F401  (the reference package)
    try:
        emu = importlib.import_module(
            "audio_processing_tools.host_analysis.device_dsd_processing_emulator")
        sys.modules["audio_processing_tools.edge.device_dsd_processing_emulator"] = emu
    except Exception:
        pass
    return root
