"""Lane-level kernel math (csrc/apt_math.cuh) compiled for the host with g++ and checked against
numpy / scipy on a CPU-only box: the same source the CUDA kernels compile."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.fft
import scipy.signal

from conftest import REPO

SRC = os.path.join(REPO, "tests", "emul", "emul.cpp")
OUT = os.path.join(REPO, "tests", "emul", "_build", "libemul.so")


@pytest.fixture(scope="module")
def emul():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [SRC, os.path.join(REPO, "audio_processing_tools_b200", "csrc", "apt_math.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-o", OUT, SRC])
    L = C.CDLL(OUT)
    L.emul_pairwise_f32.restype = C.c_float
    return L


def test_pcm_conversion_exact(emul):
    """int16 -> float32 without a divide equals float32(i)/float32(32767) for every input (audio_io.py:71-72)."""
    y = np.empty(65536, np.float32)
    emul.emul_pcm(y.ctypes.data)
    ref = np.arange(-32768, 32768).astype(np.int16).astype(np.float32) / np.float32(32767.0)
    assert np.array_equal(y, ref)


@pytest.mark.parametrize("fn,tol_frame_max,min_equal", [("emul_rfft256_f64", 1e-7, 0.999), ("emul_rfft256_f32", 2e-6, 0.0)])
def test_rfft256_against_scipy(emul, fn, tol_frame_max, min_equal):
    """The 8-lane radix-16x8 real FFT against scipy.fft.rfft in float64 rounded to complex64 (librosa's arithmetic)."""
    rng = np.random.default_rng(3)
    win = scipy.signal.get_window("hann", 256, fftbins=True)
    equal = total = 0
    for _ in range(200):
        x = (rng.standard_normal(256) * 10 ** rng.uniform(-3, 0)).astype(np.float32)
        out = np.empty(2 * 129, np.float64)
        getattr(emul, fn)(x.ctypes.data, win.ctypes.data, out.ctypes.data)
        got = (out[0::2] + 1j * out[1::2]).astype(np.complex64)
        ref = scipy.fft.rfft(win * x.astype(np.float64)).astype(np.complex64)
        assert np.abs(got - ref).max() <= tol_frame_max * np.abs(ref).max()
        equal += int((got == ref).sum())
        total += ref.size
    assert equal / total >= min_equal


def test_numpy_float32_kernels(emul):
    rng = np.random.default_rng(5)
    z = (rng.standard_normal(100_000) + 1j * rng.standard_normal(100_000)).astype(np.complex64)
    y = np.empty(z.size, np.float32)
    emul.emul_cabsf(z.ctypes.data, y.ctypes.data, C.c_long(z.size))
    assert np.array_equal(y, np.abs(z))
    for n in (1, 4, 7, 8, 9, 71, 128, 129, 256, 1000):
        a = rng.standard_normal(n).astype(np.float32)
        assert emul.emul_pairwise_f32(a.ctypes.data, n) == np.sum(a)
    x = (10 ** rng.uniform(-9, 1, 100_000)).astype(np.float32)
    emul.emul_log10f(x.ctypes.data, y.ctypes.data, C.c_long(x.size))
    assert np.max(np.abs(y.astype(np.float64) - np.log10(x.astype(np.float64))) / np.spacing(np.abs(np.log10(x)))) < 4.5
    x1 = rng.uniform(0, 60, 100_000).astype(np.float32)
    y1 = np.empty_like(x1)
    emul.emul_log1pf(x1.ctypes.data, y1.ctypes.data, C.c_long(x1.size))
    assert np.allclose(y1, np.log1p(x1), rtol=1e-6, atol=1e-7)
    try:
        from numpy._core._multiarray_umath import __cpu_features__ as feats
    except Exception:
        feats = {}
    if feats.get("AVX512_SKX"):   # numpy takes the SVML path there: bit-equal
        assert np.array_equal(y, np.log10(x))
        assert np.array_equal(y1, np.log1p(x1))


def test_power_phase_reciprocal_is_exact():
    """stft256_kernel maps element e of the (frame, bin) tile to frame e // nk with (e * (2^24 // nk + 1)) >> 24;
    exact for every nk the kernel can see (1..129 bins) and every e of a 32-frame tile."""
    for nk in range(1, 130):
        inv = (1 << 24) // nk + 1
        e = np.arange(32 * 129, dtype=np.uint64)
        assert np.array_equal((e * np.uint64(inv)) >> np.uint64(24), e // np.uint64(nk)), nk


def test_exchange_swizzle_is_conflict_free():
    """Exchange layout ex[k1 * 8 + (j ^ (k1 & 7))] (apt_math.cuh): a pass-A store (8 lanes j, one row k1) fills one
    128-byte row, and a pass-B gather (8 lanes t, rows t or 16 - t, one column j) touches 8 different 16-byte bank
    groups, so neither needs more than one shared-memory wavefront per quarter warp."""
    def idx(k1, j):
        return k1 * 8 + (j ^ (k1 & 7))
    for k1 in range(16):
        assert sorted(idx(k1, j) for j in range(8)) == list(range(k1 * 8, k1 * 8 + 8))
    for j in range(8):
        rows_a = [t for t in range(8)]                       # ka = t (lane 0 reads row 0)
        rows_b = [8 if t == 0 else 16 - t for t in range(8)]
        for rows in (rows_a, rows_b):
            groups = [(idx(r, j) * 16 // 16) % 8 for r in rows]      # 16-byte slot inside the 128-byte bank window
            assert sorted(groups) == list(range(8)), (j, rows)


def test_db_monotone_except_six_intervals(emul):
    """The median select of the noise-floor dB values runs on the bit patterns of w = N2 + eps instead of on d(w) =
    10 * log10f(w) (csrc/apt_kernels.cuh, dbsum / sel_* kernels).  That is exact as long as d is monotone in w; this scans
    EVERY float32 in [1e-9, 2^20] with the kernel's own log10 polynomial and pins the only places where it is not: six
    intervals of at most 24 floats at w = 1.5 * 2^k, which the kernels exclude (kDbExcl) by falling back to the select
    on the dB keys themselves."""
    import re
    lo = int(np.float32(1e-9).view(np.uint32))
    hi = int(np.float32(2.0 ** 20).view(np.uint32))
    out = np.zeros(2 * 64, np.uint32)
    emul.emul_db_monotone_scan.restype = C.c_int
    n = emul.emul_db_monotone_scan(C.c_uint(lo), C.c_uint(hi), out.ctypes.data, 64, os.cpu_count() or 4)
    got = sorted((int(out[2 * i]), int(out[2 * i + 1])) for i in range(n))
    src = open(os.path.join(REPO, "audio_processing_tools_b200", "csrc", "apt_kernels.cuh")).read()
    body = src[src.index("kDbExcl[6][2]"):]
    body = body[:body.index("};")]
    table = sorted((int(a, 16), int(b, 16)) for a, b in re.findall(r"\{0x([0-9a-f]{8})u, 0x([0-9a-f]{8})u\}", body))
    assert len(table) == 6
    assert got == table, (got, table)
    for a, b in got:
        assert b - a < 24
