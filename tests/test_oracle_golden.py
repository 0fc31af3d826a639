"""Pins the CPU oracle (oracle/apt_oracle.c) against golden vectors that were produced by
running the UNMODIFIED reference (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from audio_processing_tools_b200.synth import pcm_to_f32


def test_numpy_semantics_pinned(oracle_mod):
    """The transcribed numpy/SVML float32 kernels equal this host's numpy where numpy takes the
    AVX512 SVML path (the golden host).  On other hosts only the host-independent ones are checked."""
    L = oracle_mod.lib()
    rng = np.random.default_rng(0)
    z = (rng.standard_normal(200_000) + 1j * rng.standard_normal(200_000)).astype(np.complex64)
    y = np.empty(z.size, np.float32)
    L.orc_cabsf_array(z.ctypes.data, y.ctypes.data, z.size)
    assert np.array_equal(y, np.abs(z))
    for n in (1, 4, 5, 7, 8, 9, 71, 128, 129, 255, 256, 1000):
        a = rng.standard_normal(n).astype(np.float32)
        assert L.orc_np_sum_f32(a.ctypes.data, n) == np.sum(a)
    x = (10 ** rng.uniform(-9, 1, 200_000)).astype(np.float32)
    L.orc_log10f_array(x.ctypes.data, y.ctypes.data, x.size)
    ref = np.log10(x)
    # always within 4 ulp of the true value (SVML "la" bound); bit-equal on AVX512 hosts
    assert np.max(np.abs(y.astype(np.float64) - np.log10(x.astype(np.float64))) / np.spacing(np.abs(ref))) < 4.5
    x1 = rng.uniform(0, 60, 200_000).astype(np.float32)
    y1 = np.empty_like(x1)
    L.orc_log1pf_array(x1.ctypes.data, y1.ctypes.data, x1.size)
    assert np.allclose(y1, np.log1p(x1), rtol=1e-6, atol=1e-7)
    try:
        from numpy._core._multiarray_umath import __cpu_features__ as feats
    except Exception:
        feats = {}
    if feats.get("AVX512_SKX"):
        assert np.array_equal(y[: x.size], ref)
        assert np.array_equal(y1, np.log1p(x1))


def test_baseline_recursion_pinned(oracle_mod):
    g = np.load("tests/golden/baseline_k8.npz") if False else None
    import os
    from conftest import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "baseline_k8.npz"))
    x = np.ascontiguousarray(g["x"])
    out = np.empty_like(x)
    W = max(3, int(round(0.5 * (11162 / 128.0))))
    eta = float(np.clip(2.0 / max(W + 1, 2), 1e-4, 1.0))
    oracle_mod.lib().orc_baseline(x.ctypes.data, x.size, 0.2, eta, float(np.clip(1.0 - eta, 0.0, 0.9999)),
                                  1.0, out.ctypes.data)
    assert np.array_equal(out, g["baseline"])


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_golden(oracle_mod, name):
    g, meta, pcm, params = load_golden(name)
    params["keep_state_debug"] = True
    m, s = oracle_mod.run(pcm_to_f32(pcm), params)
    # events / labels: bit-exact
    assert np.array_equal(s["frame_class"], g["frame_class"])
    assert np.array_equal(s["event_idx"], g["event_idx"])
    assert np.array_equal(s["rain_conf"], g["rain_conf"])
    assert np.array_equal(s["noise_conf"], g["noise_conf"])
    assert np.array_equal(s["times"], g["times"])
    for k in ("rain_frame_count", "clip_is_rain", "clip_rain_conf", "median_rain_conf",
              "clip_rain_fraction", "clip_rain_min_frames"):
        assert m[k] == g["metric_" + k].item(), k
    # the restatement is bit-exact on this path; keep the asserted bound at 1e-6 for hosts whose
    # numpy reduces in another order
    assert m["mean_noise_floor_db"] == pytest.approx(g["metric_mean_noise_floor_db"].item(), rel=1e-6)
    assert m["median_noise_floor_db"] == pytest.approx(g["metric_median_noise_floor_db"].item(), rel=1e-6)
    if "det_primary_mode_flux" in g:
        nf = s["norm_flux"]
        for i, k in enumerate(("primary_mode_flux", "support_mode_flux_1", "support_mode_flux_2",
                               "support_mode_flux_3", "support_mode_flux_4")):
            if i < nf.shape[0]:
                assert np.array_equal(nf[i], g["det_" + k]), k
        assert np.array_equal(s["score"], g["det_mode_flux_score"])
        assert np.array_equal(s["gate"].astype(bool), g["det_td_gate_mask"])
        for i, k in enumerate(oracle_mod.TD_NAMES):
            if k == "td_kurtosis":   # numpy SVML powf inside scipy's _moment: tolerance feature
                np.testing.assert_allclose(s["td"][i], g["det_" + k], rtol=2e-6, atol=1e-6)
            else:
                assert np.array_equal(s["td"][i], g["det_" + k]), k
    if "S" in g:
        assert np.array_equal(s["S"], g["S"])
        assert np.array_equal(s["N1_band"], g["detector_noise_psd_band"])
        assert np.array_equal(s["Nlag_band"], g["detector_noise_psd_lag_band"])
        assert np.array_equal(s["N2_band"], g["noise_psd_band"])
        for i, k in enumerate(oracle_mod.RAW_NAMES):
            np.testing.assert_allclose(s["raw"][i], g["det_" + k], rtol=1e-6, atol=1e-30, err_msg=k)


def test_oracle_batch_driver_matches_single(oracle_mod):
    from audio_processing_tools_b200.synth import default_params, synth_clip_i16
    clips = [synth_clip_i16(6 + i, 100 + i, (0.0, 3.0, 10.0)[i]) for i in range(3)]
    params = default_params(check_duration=6)
    fcs, cnt = oracle_mod.process_batch_i16(clips, params, n_threads=2)
    for c, fc, n in zip(clips, fcs, cnt):
        m, s = oracle_mod.run(pcm_to_f32(c), params)
        assert np.array_equal(fc, s["frame_class"])
        assert n == m["rain_frame_count"]


SWEEP = ((256, 64), (512, 256), (1024, 256), (2048, 1024), (4096, 1024))


def load_sweep(n_fft, hop):
    import hashlib
    import json
    import os
    from conftest import GOLDEN_DIR
    from audio_processing_tools_b200.synth import default_params
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"sweep_nfft{n_fft}_hop{hop}.npz"), allow_pickle=False))
    meta = json.loads(str(g["meta"]))
    assert hashlib.sha1(g["pcm"].tobytes()).hexdigest() == meta["pcm_sha1"]
    params = default_params(check_duration=meta["seconds"], n_fft=n_fft, hop=hop)
    return g, meta, params


@pytest.mark.parametrize("n_fft,hop", SWEEP)
def test_oracle_frame_size_sweep(oracle_mod, n_fft, hop):
    """BASELINE config 5: the oracle's STFT and band-energy features at other frame sizes against the reference."""
    g, meta, params = load_sweep(n_fft, hop)
    s = oracle_mod.process(pcm_to_f32(g["pcm"]), dict(params, keep_state_debug=True))
    assert s["S"].shape == g["S"].shape
    err = np.abs(s["S"] - g["S"]).max(axis=1) / np.abs(g["S"]).max(axis=1)
    assert err.max() <= 1e-6                      # float64 FFT rounded to complex64: other algorithm, same values
    assert (s["S"] != g["S"]).mean() < 2e-3
    for i, k in enumerate(oracle_mod.RAW_NAMES):
        np.testing.assert_allclose(s["raw"][i], g["det_" + k], rtol=1e-4, atol=1e-6, err_msg=k)


GEOM = ((512, 256), (1024, 256), (2048, 1024), (4096, 1024), (256, 256), (512, 128), (256, 64), (512, 192))


@pytest.mark.parametrize("n_fft,hop", GEOM)
def test_oracle_full_pipeline_other_frame_sizes(oracle_mod, n_fft, hop):
    """The whole detector at frame sizes other than 256 / 128 (oracle/make_golden_geom.py froze the unmodified
    reference): labels, confidences, clip statistics, gate and per-mode flux, bit for bit."""
    g, meta, pcm, params = load_golden(f"geom_nfft{n_fft}_hop{hop}")
    params["keep_state_debug"] = True
    m, s = oracle_mod.run(pcm_to_f32(pcm), params)
    assert np.array_equal(s["frame_class"], g["frame_class"])
    assert np.array_equal(s["rain_conf"], g["rain_conf"])
    assert np.array_equal(s["noise_conf"], g["noise_conf"])
    assert np.array_equal(s["td"][0], g["det_td_crest_factor"])
    assert np.array_equal(s["gate"].astype(bool), g["det_td_gate_mask"])
    assert np.array_equal(s["score"], g["det_mode_flux_score"])
    assert m["rain_frame_count"] == g["metric_rain_frame_count"].item()
    assert m["mean_noise_floor_db"] == pytest.approx(g["metric_mean_noise_floor_db"].item(), rel=1e-6)
    assert m["median_noise_floor_db"] == pytest.approx(g["metric_median_noise_floor_db"].item(), rel=1e-6)
