"""Mark-3 container parsing and the batched pinned-buffer loader; CPU only."""
import struct

import numpy as np
import pytest

from audio_processing_tools_b200 import parse
from audio_processing_tools_b200.synth import synth_clip_i16


def test_header_round_trip_and_metadata():
    pcm = synth_clip_i16(2.0, 5, 3.0)
    blob = parse.build_mark_audio_file(pcm, ts=1714560000, sample_rate=11162, device_id="C012345", gps=(37.5, -122.25, 12.0))
    assert len(blob) == 40 + 2 * pcm.size and blob[:4] == b"\xAD\xFB\xCA\xDE"
    assert struct.unpack_from("<I", blob, 4)[0] == 1714560000 and blob[28:38].rstrip(b"\x00") == b"C012345"
    sig, meta = parse.parse_mark_audio_file(blob)
    assert sig.dtype == np.int16 and np.array_equal(sig, pcm)
    assert meta == {"sample_rate": 11162, "channels": 1, "bit_depth": 16, "endianness": 0, "device_id": "C012345",
                    "time": 1714560000, "lat": 37.5, "long": -122.25, "duration": round(pcm.size / 11162, 2),
                    "audio_file_version": 0, "format": "pcm"}
    assert np.array_equal(parse.pcm_to_float(sig), pcm / 32768)


def test_edge_cases():
    pcm = np.arange(-5, 6, dtype=np.int16)
    # big-endian payload, odd trailing byte dropped
    sig, meta = parse.parse_mark_audio_file(parse.build_mark_audio_file(pcm, endianness=1) + b"\x01")
    assert np.array_equal(sig, pcm) and meta["endianness"] == 1
    # no magic: the whole buffer is raw little-endian PCM with the defaults (parse.py:201-212)
    sig, meta = parse.parse_mark_audio_file(pcm.tobytes())
    assert np.array_equal(sig, pcm) and meta["sample_rate"] == 11162 and meta["device_id"] is None
    # empty payload
    sig, meta = parse.parse_mark_audio_file(parse.build_mark_audio_file(np.zeros(0, np.int16)))
    assert sig.size == 0 and meta["duration"] == 0.0
    with pytest.raises(NotImplementedError):
        parse.parse_mark_audio_file(parse.build_mark_audio_file(pcm, version=1))


def test_batch_loader_packs_contiguously():
    clips = [synth_clip_i16(1.0 + 0.37 * i, 20 + i, 3.0) for i in range(5)]
    files = [parse.build_mark_audio_file(c, ts=100 + i, device_id=f"D{i}") for i, c in enumerate(clips)]
    loader = parse.Mark3BatchLoader(sum(c.size for c in clips), pin=False)
    pcm, lengths, metas = loader.load(files)
    assert np.array_equal(pcm, np.concatenate(clips)) and list(lengths) == [c.size for c in clips]
    assert [m["device_id"] for m in metas] == [f"D{i}" for i in range(5)] and [m["time"] for m in metas] == [100 + i for i in range(5)]
    pcm, lengths, _ = loader.load(files, max_samples=11162)
    assert list(lengths) == [11162] * 5 and np.array_equal(pcm[:11162], clips[0][:11162])
    with pytest.raises(ValueError):
        parse.Mark3BatchLoader(10, pin=False).load(files)


def test_parser_matches_unmodified_reference():
    """tests/golden/mark3_cases.npz holds what the reference's own parse.parse_mark_audio_file (parse.py:164-289) returned
    for six byte strings assembled with struct.pack (oracle/make_golden_mark3.py: little / big endian, odd trailing byte,
    bit depth 0, channel flag 2, missing magic, empty payload, forced PCM on a version-1 header): samples, dtype, every
    metadata field and pcm_to_float must be identical."""
    import json
    import os
    from conftest import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "mark3_cases.npz"), allow_pickle=False)
    index = json.loads(str(g["index"]))
    assert len(index) == 6
    for case in index:
        name = case["name"]
        blob = g[f"{name}__blob"].tobytes()
        sig, meta = parse.parse_mark_audio_file(blob, force_file_type=case["force"])
        assert str(sig.dtype) == case["sig_dtype"], name
        assert np.array_equal(sig, g[f"{name}__sig"]), name
        assert meta == case["meta"], (name, meta, case["meta"])
        assert np.array_equal(parse.pcm_to_float(sig), g[f"{name}__float"]), name
