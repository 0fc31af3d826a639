#!/usr/bin/env python
"""Golden vectors for the Mark-3 input path by RUNNING THE UNMODIFIED REFERENCE parser.

`parse.parse_mark_audio_file` (reference parse.py:164-289, Kaitai class AudioBinary :29-54, _decode_pcm_payload :539-580,
pcm_to_float :670) is imported through oracle/refharness (its kaitaistruct stand-in restates the five stream reads the
generated class makes) and run on byte strings assembled HERE with struct.pack -- not with the product's own writer --
so that tests/test_parse_mark3.py pins the product's parser to the reference's, field by field.

    python oracle/make_golden_mark3.py   ->  tests/golden/mark3_cases.npz
"""
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)
import refharness  # noqa: E402

refharness.install()
from audio_processing_tools import parse as ref_parse  # noqa: E402


def header(ts, sr, ch, depth, endian, version, gps, dev, pad=b"\x00\x00"):
    return (b"\xAD\xFB\xCA\xDE" + struct.pack("<II", ts, sr) + struct.pack("<BBBB", ch, depth, endian, version) +
            struct.pack("<fff", *gps) + dev.encode("utf-8").ljust(10, b"\x00")[:10] + pad)


def main():
    rng = np.random.default_rng(42)
    cases = []
    pcm_a = rng.integers(-32768, 32767, 4321, dtype=np.int16)
    pcm_b = rng.integers(-2000, 2000, 1000, dtype=np.int16)
    pcm_c = rng.integers(-32768, 32767, 777, dtype=np.int16)
    cases.append(("le_mono", header(1714560000, 11162, 1, 16, 0, 0, (37.5, -122.25, 12.0), "C012345") + pcm_a.astype("<i2").tobytes(), None))
    cases.append(("be_odd_tail", header(1700000001, 11162, 1, 16, 1, 0, (0.0, 0.0, 0.0), "DEVICE0001") + pcm_b.astype(">i2").tobytes() + b"\x7f", None))
    cases.append(("depth0_stereo_flag", header(5, 8000, 2, 0, 0, 0, (-33.125, 151.5, -3.0), "X") + pcm_c.astype("<i2").tobytes()[:-2], None))
    cases.append(("no_magic_raw", pcm_b.astype("<i2").tobytes(), None))
    cases.append(("empty_payload", header(9, 11162, 1, 16, 0, 0, (1.0, 2.0, 3.0), "E"), None))
    cases.append(("forced_pcm_v1", header(77, 11162, 1, 16, 0, 1, (0.5, 0.25, 0.0), "ALACDEV") + pcm_a[:100].astype("<i2").tobytes(), "pcm"))
    out = {}
    index = []
    for name, blob, force in cases:
        sig, meta = ref_parse.parse_mark_audio_file(blob, force_file_type=force)
        out[f"{name}__blob"] = np.frombuffer(blob, dtype=np.uint8)
        out[f"{name}__sig"] = np.asarray(sig)
        out[f"{name}__float"] = np.asarray(ref_parse.pcm_to_float(np.asarray(sig)))
        meta = {k: (None if v is None else (float(v) if isinstance(v, float) else v)) for k, v in meta.items()}
        index.append({"name": name, "force": force, "meta": meta, "sig_dtype": str(np.asarray(sig).dtype)})
        print(name, len(blob), np.asarray(sig).shape, meta)
    out["index"] = np.array(json.dumps(index))
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "mark3_cases.npz"), **out)


if __name__ == "__main__":
    main()
