"""ctypes binding of libapt_b200.so (include/apt_b200.h).

The shared library is the product: there is no CPU fallback.  If it is missing or no CUDA
device is visible every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import glob
import hashlib
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "libapt_b200.so")
CSRC = os.path.join(_PKG, "csrc")

MAX_MODES, MAX_SOS, N_RAW, N_TD, N_STATS = 8, 4, 21, 5, 8
ABI_VERSION = 10
MAX_GAIN_TAPS = 9
MAX_PRE_SMOOTH, MAX_MEDIAN = 16, 31
STAGE_FEATURES, STAGE_FULL = 1, 2
KERNEL_NAMES = ("stft256_kernel", "td_features_kernel", "trk1_kernel", "flux_kernel", "base_kernel",
                "decide_kernels", "trk2_kernel", "db_kernel", "select_kernels", "finalize_kernel", "gain_kernels")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false",
              "-std=c++17", "-shared", "-Xcompiler", "-fPIC"]


class AptParams(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("fs", C.c_int32), ("n_fft", C.c_int32), ("hop", C.c_int32),
        ("band_lo", C.c_int32), ("band_hi", C.c_int32), ("n_modes", C.c_int32),
        ("mode_lo", C.c_int32 * MAX_MODES), ("mode_hi", C.c_int32 * MAX_MODES),
        ("mode_band_lo", C.c_int32 * MAX_MODES), ("mode_band_hi", C.c_int32 * MAX_MODES),
        ("mode_weight", C.c_double * MAX_MODES),
        ("trk_eta", C.c_float), ("trk_scale_alpha", C.c_float), ("trk_one_minus_alpha", C.c_float),
        ("trk_step_floor", C.c_float), ("trk_q", C.c_float), ("trk_neg_one_minus_q", C.c_float),
        ("trk_maxr", C.c_float),
        ("ema_up", C.c_double), ("ema_down", C.c_double),
        ("warmup_need", C.c_int32), ("eps_f32", C.c_float),
        ("detector_use_noise_norm", C.c_int32), ("norm_ratio_db", C.c_int32),
        ("bl_q", C.c_double), ("bl_eta", C.c_double), ("bl_scale_alpha", C.c_double), ("bl_floor", C.c_double),
        ("norm_enable", C.c_int32), ("norm_min_f32", C.c_float),
        ("thr_primary", C.c_float), ("thr_m1", C.c_float), ("thr_m2", C.c_float), ("thr_m3", C.c_float),
        ("min_support", C.c_int32), ("td_gate_thr", C.c_float), ("has_kurt_upper", C.c_int32),
        ("kurt_upper", C.c_float), ("noise_hi", C.c_float), ("mode_flux_noise_max", C.c_float),
        ("n_sos", C.c_int32), ("padlen", C.c_int32),
        ("sos", (C.c_double * 6) * MAX_SOS), ("zi", (C.c_double * 2) * MAX_SOS),
        ("eps_f64", C.c_double),
        ("blk_len", C.c_int32), ("blk_hop", C.c_int32), ("blk_post_pre", C.c_int32), ("blk_smooth", C.c_int32),
        ("low_lo", C.c_int32), ("low_hi", C.c_int32), ("rain_lo", C.c_int32), ("rain_hi", C.c_int32),
        ("rolloff_fraction", C.c_double),
        ("suppressor_bypass", C.c_int32), ("clip_rain_min_frames", C.c_int32),
        ("fft_f64", C.c_int32),
        ("gain_mode", C.c_int32), ("adaptive_gain", C.c_int32), ("gain_freq_smooth", C.c_int32),
        ("n_gain_taps", C.c_int32), ("use_lagged_noise_psd", C.c_int32),
        ("oversub_noise", C.c_float), ("oversub_rain", C.c_float), ("gain_floor", C.c_float), ("gain_ceil", C.c_float),
        ("gain_taps", C.c_float * MAX_GAIN_TAPS),
        ("alpha_noise", C.c_float), ("one_minus_alpha_noise", C.c_float),
        ("alpha_base", C.c_float), ("one_minus_alpha_base", C.c_float), ("gain_eps_f32", C.c_float),
        ("peak_top_p", C.c_int32), ("primary_top_m", C.c_int32),
        ("peak_prominence_db", C.c_double), ("peak_min_db_above_floor", C.c_double), ("peak_ratio_min", C.c_double),
        ("peak_valid_prom_min_db", C.c_float), ("peak_valid_prom_max_db", C.c_float),
        ("adaptive_q", C.c_int32), ("pre_smooth_frames", C.c_int32),
        ("aq_base", C.c_double), ("aq_min", C.c_double), ("aq_alpha", C.c_double),
        ("median_frames", C.c_int32), ("snr_gating", C.c_int32), ("snr_gating_snr1", C.c_float),
        ("snr_gating_power", C.c_float),
        ("snr_mask", C.c_uint32 * 4), ("bypass_classifier", C.c_int32),
        ("window", C.c_void_p), ("freqs", C.c_void_p),
    ]


OUT_FIELDS = ("frame_class", "rain_conf", "noise_conf", "event_idx", "event_count", "clip_stats",
              "S", "P", "det_noise_psd", "det_noise_lag", "D", "noise_psd", "mode_flux", "norm_flux",
              "score", "td", "raw", "band_energy", "gate", "x_td", "G", "ratio_med", "S_hat",
              "peak_ratio", "peak_gate_score", "peak_valid_count", "peak_count_by_mode", "y", "td_fast_crest",
              "snr_mode", "snr_gate")


class AptDsdParams(C.Structure):
    _fields_ = [("fs", C.c_int32), ("frame_length", C.c_int32), ("hop_length", C.c_int32), ("apply_window", C.c_int32),
                ("window", C.c_void_p)]


BNE_MAX_SOS, BNE_MAX_S, BNE_MAX_BANDS, BNE_FRAME_F, BNE_STATS = 8, 8, 8, 12, 16


class AptBneParams(C.Structure):
    _fields_ = [
        ("fs", C.c_int32), ("N", C.c_int32), ("sub_len", C.c_int32), ("S", C.c_int32),
        ("ns_h", C.c_int32), ("ns_b", C.c_int32), ("warm", C.c_int32), ("pad0", C.c_int32),
        ("sos_h", (C.c_double * 6) * BNE_MAX_SOS), ("zi_h", (C.c_double * 2) * BNE_MAX_SOS),
        ("sos_b", (C.c_double * 6) * BNE_MAX_SOS), ("zi_b", (C.c_double * 2) * BNE_MAX_SOS),
        ("n_bands", C.c_int32), ("band_b0", C.c_int32 * BNE_MAX_BANDS), ("band_b1", C.c_int32 * BNE_MAX_BANDS),
        ("prim_b0", C.c_int32), ("prim_b1", C.c_int32), ("mask_b0", C.c_int32), ("mask_b1", C.c_int32), ("pad1", C.c_int32),
        ("M_ratio", C.c_double), ("N_ratio", C.c_double), ("D_ratio", C.c_double), ("band_rise_db", C.c_double),
        ("excess_rise_db", C.c_double), ("min_Ehpf", C.c_double), ("min_Eband", C.c_double), ("dE_thr", C.c_double),
        ("k_subframes", C.c_int32), ("use_dE", C.c_int32), ("use_D", C.c_int32), ("pad2", C.c_int32),
        ("W", C.c_int32), ("W_min", C.c_int32), ("ttl", C.c_int32), ("smooth", C.c_int32), ("learn_all", C.c_int32),
        ("replenish", C.c_int32), ("replenish_only_not_full", C.c_int32), ("q_adapt", C.c_int32),
        ("q", C.c_double), ("ema_alpha", C.c_double), ("beta", C.c_double), ("gain_floor", C.c_double), ("eps", C.c_double),
        ("att_dry", C.c_double), ("att_wet", C.c_double), ("release", C.c_double), ("repl_q", C.c_double),
        ("q_repl_alpha", C.c_double), ("q_norm_alpha", C.c_double),
    ]


ROE_FRAME_F, ROE_PART_F, ROE_CLIP_F = 8, 5, 5


class AptRoeParams(C.Structure):
    _fields_ = [
        ("n_fft", C.c_int32), ("hop", C.c_int32), ("M", C.c_int32), ("wl", C.c_int32), ("max_peaks", C.c_int32), ("want_td", C.c_int32),
        ("ns_in", C.c_int32), ("ns_td", C.c_int32),
        ("sos_in", (C.c_double * 6) * 8), ("sos_td", (C.c_double * 6) * 4), ("window", C.c_double * 256), ("fs", C.c_double),
        ("f_natural", C.c_double), ("op_lo", C.c_double), ("op_hi", C.c_double), ("nat_lo", C.c_double), ("nat_hi", C.c_double),
        ("search0_lo", C.c_double), ("search0_hi", C.c_double), ("rain_thr", C.c_double * 6), ("rain_thr_hn", C.c_double),
        ("kurtosis_thr", C.c_double), ("crest_thr", C.c_double), ("diff_energy_thr", C.c_double),
        ("handle_fp", C.c_int32), ("handle_fn", C.c_int32), ("rain_drop_threshold", C.c_int32), ("rain_drop_max_thr", C.c_int32),
        ("rain_peaks_min_thr", C.c_int32), ("rain_peaks_max_thr", C.c_int32),
    ]


class AptOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in OUT_FIELDS]


# every symbol include/apt_b200.h declares
EXPORTS = ("apt_init", "apt_destroy", "apt_last_error", "apt_abi_version", "apt_sizeof_params",
           "apt_sizeof_out", "apt_params_default", "apt_plan_create", "apt_plan_destroy",
           "apt_plan_offsets", "apt_plan_total_frames", "apt_plan_total_samples",
           "apt_plan_scratch_bytes", "apt_run_i16", "apt_run_f32", "apt_plan_last_launches",
           "apt_run_host_i16", "apt_run_host_clips", "apt_source_hash", "apt_plan_enable_timing", "apt_plan_enable_trace", "apt_plan_trace", "apt_plan_tc_error", "apt_plan_kernel_ms", "apt_selftest",
           "apt_dsd_run_i16", "apt_sizeof_bne_params", "apt_bne_run",
           "apt_sizeof_roe_params", "apt_roe_run")


def source_files():
    """Every file the library is compiled from: all of csrc/ plus the public header(s)."""
    files = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                   glob.glob(os.path.join(_REPO, "include", "*.h")))
    return files


def source_hash():
    """sha256 over the sources and the compiler flags; compiled into the library (apt_source_hash) so that a
    stale binary is detected by content, not by file times (the .so is git-ignored and travels prebuilt)."""
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for f in source_files():
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:32]


def built_hash():
    """Hash recorded in the built library, or None.  Read from the file's bytes (the library carries the string
    "APT_SRC_HASH=<hash>"), not through dlopen: a library loaded once stays mapped under its path for the life of
    the process, so a rebuild could not be checked that way."""
    if not os.path.exists(LIB_PATH):
        return None
    data = open(LIB_PATH, "rb").read()
    i = data.find(b"APT_SRC_HASH=")
    if i < 0:
        return None
    return data[i + 13:i + 13 + 32].decode("ascii", "replace")


def build(force=False, verbose=False):
    """Compile csrc/apt_b200.cu for sm_100a into the package directory (in-tree .so).  Rebuilds whenever the
    hash of the sources differs from the one compiled into the existing library."""
    want = source_hash()
    if not force and built_hash() == want:
        return LIB_PATH
    tmp = LIB_PATH + ".tmp"
    cmd = ["nvcc"] + NVCC_FLAGS + ["-DAPT_SRC_HASH=\"%s\"" % want, "-o", tmp, os.path.join(CSRC, "apt_b200.cu")]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd, cwd=CSRC)
    os.replace(tmp, LIB_PATH)
    global _lib
    _lib = None
    return LIB_PATH


_lib = None


def load():
    """dlopen the library and type its entry points.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`."
            " There is no CPU fallback for this path.")
    L = C.CDLL(LIB_PATH)
    L.apt_source_hash.restype = C.c_char_p
    if os.path.isdir(CSRC) and L.apt_source_hash().decode() != source_hash():
        raise RuntimeError(
            f"{LIB_PATH} was built from other sources than the ones in {CSRC} (hash "
            f"{L.apt_source_hash().decode()} != {source_hash()}): rebuild with "
            "`python -c 'import __graft_entry__ as g; g.build()'`")
    vp, i64p = C.c_void_p, C.POINTER(C.c_int64)
    L.apt_init.argtypes = [C.c_int, C.POINTER(vp)]
    L.apt_destroy.argtypes = [vp]
    L.apt_last_error.argtypes = [vp]
    L.apt_last_error.restype = C.c_char_p
    L.apt_params_default.argtypes = [C.POINTER(AptParams)]
    L.apt_plan_create.argtypes = [vp, C.POINTER(AptParams), C.c_int, i64p, C.POINTER(vp)]
    L.apt_plan_destroy.argtypes = [vp]
    L.apt_plan_offsets.argtypes = [vp, i64p, i64p]
    for f in ("apt_plan_total_frames", "apt_plan_total_samples", "apt_plan_scratch_bytes"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = C.c_int64
    L.apt_plan_last_launches.argtypes = [vp]
    L.apt_run_i16.argtypes = [vp, C.c_int, vp, C.POINTER(AptOut), vp]
    L.apt_run_f32.argtypes = [vp, C.c_int, vp, C.POINTER(AptOut), vp]
    L.apt_run_host_i16.argtypes = [vp] * 8
    L.apt_run_host_clips.argtypes = [vp, C.POINTER(vp), C.c_int] + [vp] * 6
    L.apt_selftest.argtypes = [vp, C.c_int, C.c_int64, i64p]
    L.apt_bne_run.argtypes = [vp, C.POINTER(AptBneParams), C.c_int, i64p, vp, C.c_int, vp, vp, vp, vp, vp]
    i32p = C.POINTER(C.c_int32)
    L.apt_roe_run.argtypes = [vp, C.POINTER(AptRoeParams), C.c_int, vp, C.c_int, C.c_int, i32p, i64p, i32p, C.c_int, vp, vp, vp,
                              C.POINTER(C.c_int), vp]
    if L.apt_sizeof_roe_params() != C.sizeof(AptRoeParams):
        raise RuntimeError("libapt_b200.so apt_roe_params_t layout differs from the Python binding")
    if L.apt_sizeof_bne_params() != C.sizeof(AptBneParams):
        raise RuntimeError("libapt_b200.so apt_bne_params_t layout differs from the Python binding")
    L.apt_dsd_run_i16.argtypes = [vp, C.POINTER(AptDsdParams), C.c_int, i64p, C.POINTER(C.c_double), vp, vp, vp, C.c_int, vp]
    L.apt_plan_enable_timing.argtypes = [vp, C.c_int]
    L.apt_plan_tc_error.argtypes = [vp]
    L.apt_plan_enable_trace.argtypes = [vp, C.c_int]
    L.apt_plan_trace.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_float)]
    L.apt_plan_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    if L.apt_abi_version() != ABI_VERSION:
        raise RuntimeError("libapt_b200.so ABI version mismatch")
    if L.apt_sizeof_params() != C.sizeof(AptParams) or L.apt_sizeof_out() != C.sizeof(AptOut):
        raise RuntimeError("libapt_b200.so struct layout differs from the Python binding")
    _lib = L
    return L


# Optional device timing of the C-ABI calls of the engines beside the main path (bench.py --workload bne|roe|dsd):
# CUDA events on the stream the call launches on, recorded around the call (inputs already in HBM).
DEVICE_TIMING = False
LAST_DEVICE_MS = {}


class device_timer:
    def __init__(self, torch, key, device):
        self.torch, self.key, self.device = torch, key, device
        self.on = DEVICE_TIMING

    def __enter__(self):
        if self.on:
            self.e0 = self.torch.cuda.Event(enable_timing=True)
            self.e1 = self.torch.cuda.Event(enable_timing=True)
            self.e0.record(self.torch.cuda.current_stream(self.device))
        return self

    def __exit__(self, *exc):
        if self.on and exc[0] is None:
            self.e1.record(self.torch.cuda.current_stream(self.device))
            self.e1.synchronize()
            LAST_DEVICE_MS[self.key] = float(self.e0.elapsed_time(self.e1))
        return False
