"""Mark-3 audio container: header + int16 PCM, straight into pinned host memory.

Mirrors the PCM side of the reference's parse.py: the 40-byte header read by the Kaitai class ``AudioBinary``
(:29-54: magic AD FB CA DE, u32 timestamp, u32 sample_rate, u8 channels / bit depth / endianness / file
version, 3 x f32 GPS, 10-byte device id, 2 pad bytes), ``parse_mark_audio_file`` (:164-289, same metadata
keys and the same fall-back to raw-PCM defaults when the magic is missing), ``_decode_pcm_payload``
(:539-580) and ``pcm_to_float`` (:670).  ALAC payloads (file version >= 1) are decoded by an ffmpeg
subprocess in the reference (:373-536); that stays host-side work outside this package and is refused here.

``Mark3BatchLoader`` is the piece the GPU path adds: it lays the PCM of many files end to end in ONE pinned
int16 buffer (the layout ``apt_run_host_i16`` consumes), so the host->device copies run at full PCIe rate and
no float conversion happens on the CPU (the device converts int16 on the fly).
"""
from __future__ import annotations

import struct
from typing import Any, Dict, Iterable, List, Optional, Tuple

import numpy as np

MAGIC = b"\xAD\xFB\xCA\xDE"
HEADER_SIZE = 40
_HDR = struct.Struct("<4sIIBBBBfff10s2s")
assert _HDR.size == HEADER_SIZE
bytes_per_sample = 2


class ValidationNotEqualError(Exception):
    """Raised when the magic bytes do not match (name kept from kaitaistruct, parse.py:19-21)."""


def parse_header(file_contents: bytes) -> Dict[str, Any]:
    if len(file_contents) < HEADER_SIZE or file_contents[:4] != MAGIC:
        raise ValidationNotEqualError(f"expected magic {MAGIC!r}, got {bytes(file_contents[:4])!r}")
    magic, ts, sr, ch, depth, endian, version, lat, lon, alt, dev, _pad = _HDR.unpack_from(file_contents, 0)
    return {"device": dev.decode("UTF-8").rstrip("\x00"), "ts": ts, "sample_rate": sr, "channels": ch,
            "bit_depth": depth, "endianness": endian, "gps": [lat, lon, alt], "audio_file_version": version,
            "audio": memoryview(file_contents)[HEADER_SIZE:]}


def build_mark_audio_file(pcm: np.ndarray, *, ts: int = 0, sample_rate: int = 11162, device_id: str = "C000000",
                          gps=(0.0, 0.0, 0.0), version: int = 0, endianness: int = 0) -> bytes:
    """Inverse of parse_mark_audio_file for PCM files (test and tooling helper)."""
    pcm = np.asarray(pcm, dtype=np.int16)
    hdr = _HDR.pack(MAGIC, int(ts), int(sample_rate), 1, 16, int(endianness), int(version),
                    float(gps[0]), float(gps[1]), float(gps[2]), device_id.encode("UTF-8")[:10], b"\x00\x00")
    return hdr + pcm.astype(">i2" if endianness else "<i2").tobytes()


def _decode_pcm_payload(audio_data, bit_depth: int, channels: int, endianness: int) -> np.ndarray:
    if bit_depth != 16:
        raise ValueError(f"Unsupported PCM bit depth: {bit_depth}")
    sig = np.frombuffer(audio_data, dtype="<i2" if endianness == 0 else ">i2")
    return sig.astype(np.int16, copy=False)


def parse_mark_audio_file(file_contents: bytes, force_file_type: Optional[str] = None) -> Tuple[np.ndarray, Dict[str, Any]]:
    try:
        h = parse_header(file_contents)
        sample_rate, channels, bit_depth, endianness = h["sample_rate"], h["channels"], h["bit_depth"], h["endianness"]
        gps, audio_data, device_id, time, file_version = h["gps"], h["audio"], h["device"], h["ts"], h["audio_file_version"]
    except ValidationNotEqualError:
        print("WARNING: Could not parse header, assuming raw PCM defaults")
        sample_rate, channels, bit_depth, endianness, file_version = 11162, 1, 16, 0, 0
        gps, device_id, time, audio_data = (None, None), None, None, memoryview(file_contents)
    if bit_depth == 0:
        bit_depth = 16
    if bit_depth % 8 != 0:
        raise ValueError(f"Invalid bit depth {bit_depth}: must be multiple of 8")
    if bit_depth != 16:
        print(f"WARNING: Unsupported bit depth {bit_depth}; assuming 16-bit PCM compatibility")
    bps = bit_depth // 8
    rem = len(audio_data) % bps
    if rem:
        audio_data = audio_data[: len(audio_data) - rem]
    is_alac = (force_file_type == "alac") or (force_file_type != "pcm" and file_version >= 1)
    if is_alac:
        raise NotImplementedError("ALAC payloads (audio_file_version >= 1) are decoded through ffmpeg in the reference "
                                  "(parse.py:373-536); decode them on the host and pass the int16 PCM")
    sig = _decode_pcm_payload(audio_data, bit_depth=bit_depth, channels=channels, endianness=endianness)
    n_per_channel = len(sig) / channels if channels > 0 else len(sig)
    metadata = {"sample_rate": sample_rate, "channels": channels, "bit_depth": bit_depth, "endianness": endianness,
                "device_id": device_id, "time": time, "lat": gps[0], "long": gps[1],
                "duration": round(n_per_channel / sample_rate, 2), "audio_file_version": file_version, "format": "pcm"}
    return sig, metadata


def pcm_to_float(signal, scale_factor=1 << (bytes_per_sample * 8 - 1)):
    return signal / scale_factor


class Mark3BatchLoader:
    """Parses Mark-3 PCM files and packs their samples into one pinned int16 buffer.

    ``load(files)`` -> (pcm: 1-D int16 numpy view of pinned memory, lengths: int64[n], metadata: list of dicts).
    ``max_samples`` trims every clip (the reference trims to sample_rate * check_duration, audio_io.py:99-120)."""

    def __init__(self, capacity_samples: int, pin: bool = True):
        self.capacity = int(capacity_samples)
        self._torch_buf = None
        if pin:
            try:
                import torch
                if torch.cuda.is_available():
                    self._torch_buf = torch.empty(self.capacity, dtype=torch.int16, pin_memory=True)
            except Exception:
                self._torch_buf = None
        self.buffer = self._torch_buf.numpy() if self._torch_buf is not None else np.empty(self.capacity, np.int16)
        self.pinned = self._torch_buf is not None

    def load(self, files: Iterable[bytes], max_samples: Optional[int] = None):
        lengths: List[int] = []
        metas: List[Dict[str, Any]] = []
        pos = 0
        for contents in files:
            sig, meta = parse_mark_audio_file(contents)
            if max_samples is not None:
                sig = sig[:max_samples]
            if pos + sig.size > self.capacity:
                raise ValueError(f"batch does not fit the loader's {self.capacity}-sample buffer")
            self.buffer[pos:pos + sig.size] = sig       # the only copy: file bytes -> pinned staging
            pos += sig.size
            lengths.append(int(sig.size))
            metas.append(meta)
        return self.buffer[:pos], np.asarray(lengths, dtype=np.int64), metas
