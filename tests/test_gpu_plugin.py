"""GPU tests of the plugin boundary: the processors a reference user instantiates (RainDetectorProcessor,
NoiseProcessor) and the batch orchestrator that calls them, on the default-flags path that goes through
apt_run_host_clips (pinned staging ring, pipelined H2D / compute / D2H), against the goldens the unmodified
reference produced (tests/golden, oracle/make_golden.py).

Reference seams: processor contract audio_processing_framework.py:52-100 / :183-207, NoiseProcessor's intended
results noise_processor.py:104-127, injected loaders audio_processing_framework.py:653-656."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from audio_processing_tools_b200.synth import FS, default_params, pcm_to_f32, synth_clip_i16

pytestmark = pytest.mark.gpu

DEFAULT_GOLDENS = [n for n in golden_names(max_seconds=60) if not n.startswith("alt_")]


@pytest.fixture(scope="module")
def goldens():
    """(golden arrays, meta, int16 PCM, params) of the reference fixtures up to 60 s, default parameters."""
    out = {}
    for name in DEFAULT_GOLDENS:
        g, meta, pcm, params = load_golden(name)
        out[name] = (g, meta, pcm, params)
    return out


def _check_detector_result(m, s, g):
    assert np.array_equal(s["frame_class"], g["frame_class"])
    assert np.array_equal(s["rain_conf"], g["rain_conf"])
    assert np.array_equal(s["noise_conf"], g["noise_conf"])
    assert np.array_equal(s["times"], g["times"])
    for k in ("rain_frame_count", "clip_is_rain", "clip_rain_conf", "median_rain_conf", "clip_rain_fraction",
              "clip_rain_min_frames"):
        assert m[k] == g["metric_" + k].item(), k
    assert m["rain_frame_fraction"] == m["clip_rain_fraction"]


def test_detector_run_batch_default_flags_ragged_int16_and_float32(goldens):
    """RainDetectorProcessor.run_batch under default flags (host path): every golden clip in ONE ragged batch, once as
    the int16 wire samples and once as the float32 waveform the reference's loader hands out."""
    from audio_processing_tools_b200.edge.rain_signal_processor import RainDetectorProcessor
    names = list(goldens)
    params = dict(goldens[names[0]][3], check_duration=6)
    proc = RainDetectorProcessor()
    for conv in (lambda p: p, pcm_to_f32):
        clips = [conv(goldens[n][2]) for n in names]
        outs = proc.run_batch(clips, params)
        assert len(outs) == len(names)
        for n, (m, s) in zip(names, outs):
            _check_detector_result(m, s, goldens[n][0])
            assert "mean_noise_floor_db" not in m          # only with noise_psd returned (rain_signal_processor.py:1286)
            assert s["processor"] == "rain_detector" and s["features"] is None
            assert s["frame_class"].dtype == np.int8 and s["rain_conf"].dtype == np.float32
    # results are caller-owned: a second batch must not overwrite the arrays handed out by the first
    keep = [s["frame_class"].copy() for _, s in outs]
    proc.run_batch([pcm_to_f32(goldens[names[-1]][2])] * 3, params)
    for k, (_, s) in zip(keep, outs):
        assert np.array_equal(k, s["frame_class"])


def test_detector_host_path_equals_planes_path(goldens):
    """The default-flags host path and the keep_state_debug path (device-resident, optional planes) agree bit for bit."""
    from audio_processing_tools_b200.edge.rain_signal_processor import RainDetectorProcessor
    names = list(goldens)[:4]
    params = dict(goldens[names[0]][3], check_duration=6)
    proc = RainDetectorProcessor()
    clips = [pcm_to_f32(goldens[n][2]) for n in names]
    fast = proc.run_batch(clips, params)
    full = proc.run_batch(clips, dict(params, keep_state_debug=True))
    for (m0, s0), (m1, s1) in zip(fast, full):
        assert np.array_equal(s0["frame_class"], s1["frame_class"])
        assert np.array_equal(s0["rain_conf"], s1["rain_conf"])
        for k in ("rain_frame_count", "clip_is_rain", "clip_rain_conf", "clip_rain_fraction"):
            assert m0[k] == m1[k]


@pytest.mark.parametrize("name", DEFAULT_GOLDENS[:6])
def test_noise_processor_run_matches_reference_metrics(goldens, name):
    """NoiseProcessor.run (intended contract, noise_processor.py:104-127): the noise-floor statistics and the rain-frame
    fraction equal what the reference's engine produced for the same clip (RainDetectorProcessor(keep_state_debug=True)
    goldens: same noise_psd); default flags = host path, no plane leaves the GPU."""
    from audio_processing_tools_b200.noise_processor import NoiseProcessor
    g, meta, pcm, params = goldens[name]
    proc = NoiseProcessor(name="noise")
    m, s = proc.run(pcm_to_f32(pcm), params)
    assert set(m) == {"mean_noise_floor_db", "median_noise_floor_db", "rain_frame_fraction", "latency_s"}
    assert m["mean_noise_floor_db"] == pytest.approx(float(g["metric_mean_noise_floor_db"]), rel=1e-5)
    assert m["median_noise_floor_db"] == pytest.approx(float(g["metric_median_noise_floor_db"]), rel=1e-6)
    assert m["rain_frame_fraction"] == float(np.mean(g["frame_class"] == 2))
    assert np.array_equal(s["is_rain"], g["frame_class"] == 2)
    assert np.array_equal(s["times"], g["times"])
    assert s["processor"] == "noise" and "noise_psd" not in s


def test_noise_processor_batch_and_full_state(goldens):
    """run_batch over a ragged batch; keep_state_full returns the arrays the reference's state lists."""
    from audio_processing_tools_b200.noise_processor import NoiseProcessor
    names = [n for n in goldens if n.startswith("full_")]
    params = dict(goldens[names[0]][3], check_duration=6)
    proc = NoiseProcessor(name="noise")
    clips = [pcm_to_f32(goldens[n][2]) for n in names]
    lean = proc.run_batch(clips, params)
    rich = proc.run_batch(clips, dict(params, keep_state_full=True))
    for n, (m0, s0), (m1, s1) in zip(names, lean, rich):
        g = goldens[n][0]
        for k in ("mean_noise_floor_db", "median_noise_floor_db", "rain_frame_fraction"):
            assert m0[k] == m1[k], k
        assert m0["median_noise_floor_db"] == pytest.approx(float(g["metric_median_noise_floor_db"]), rel=1e-6)
        band = s1["noise_psd"][10:81].T
        np.testing.assert_allclose(band, g["noise_psd_band"], rtol=1e-4, atol=1e-12)
        assert s1["noise_psd"].shape == (129, g["frame_class"].size)
        # the statistics are the reference's formulas applied to the returned plane (noise_processor.py:104-112)
        db = 10.0 * np.log10(s1["noise_psd"][10:81] + np.float32(1e-9))
        assert m1["mean_noise_floor_db"] == pytest.approx(float(np.mean(db)), rel=1e-5)
        assert m1["median_noise_floor_db"] == pytest.approx(float(np.median(db)), rel=1e-6)
        T = g["frame_class"].size
        assert s1["S"].shape == (129, T) and s1["S_hat"].shape == (129, T)
        assert s1["input_audio"].shape == clips[names.index(n)].shape
        assert s1["denoised_audio"].shape == clips[names.index(n)].shape
        assert "debug" in s1 and "config" in s1
    with pytest.raises(ValueError):
        proc.run(clips[0], dict(params, suppressor_bypass=True))
    with pytest.raises(ValueError):
        proc.run(clips[0][:100], params)


def test_orchestrator_with_gpu_processors(goldens, tmp_path):
    """process_audio_batches_v2 with the real GPU processors and an injected synthetic loader: rows equal the goldens;
    batch_size smaller than the corpus; one file too short (dropped before run, framework :162-169)."""
    import pandas as pd
    from audio_processing_tools_b200.audio_processing_framework import process_audio_batches_v2
    from audio_processing_tools_b200.edge.rain_signal_processor import RainDetectorProcessor
    from audio_processing_tools_b200.noise_processor import NoiseProcessor
    names = [n for n in goldens if goldens[n][1]["seconds"] == 60]
    assert len(names) >= 5
    base = goldens[names[0]][3]

    def get_keys(InputType, **kw):
        return [{"source_file": n, "raining": goldens[n][1]["lam"] > 0} for n in names] + \
               [{"source_file": "zz_short", "raining": False}]

    def loader(keys, InputType, Fs, check_duration, localStatus, local_cache, read_size=None, bytes_per_sample=2, **kw):
        out = {}
        for k in keys:
            n = k["source_file"]
            audio = np.zeros(1000, np.float32) if n == "zz_short" else pcm_to_f32(goldens[n][2])
            out[n] = {"file_contents": audio, "raining": k["raining"]}
        return out

    procs = [RainDetectorProcessor(name="rain_detector"), NoiseProcessor(name="noise")]
    res, states = process_audio_batches_v2(
        processors=procs, params_global={"sample_rate": FS, "check_duration": 60, "detector": base["detector"]},
        batch_size=2, batch_save_dir=str(tmp_path), max_batch_save=3, get_keys_fn=get_keys, get_input_data_fn=loader)
    # rows are flushed to parquet every >= 3 rows (framework :813-829) and the in-memory frames keep the remainder:
    # the parts together hold every file once, without the short one (dropped before run)
    parts = res.attrs["saved_parquet_files"]
    assert len(parts) == 2
    rows = pd.concat([pd.read_parquet(p) for p in parts]).sort_values("file_key").reset_index(drop=True)
    assert list(rows["file_key"]) == sorted(names)
    assert list(res["file_key"]) == names[4:]               # the keys arrive in the loader's order, not sorted
    for _, row in rows.iterrows():
        g = goldens[row["file_key"]][0]
        assert row["rain_detector__rain_frame_count"] == int(g["metric_rain_frame_count"])
        assert bool(row["rain_detector__clip_is_rain"]) == bool(g["metric_clip_is_rain"])
        assert row["rain_detector__clip_rain_fraction"] == float(g["metric_clip_rain_fraction"])
        assert row["noise__median_noise_floor_db"] == pytest.approx(float(g["metric_median_noise_floor_db"]), rel=1e-6)
        assert row["noise__mean_noise_floor_db"] == pytest.approx(float(g["metric_mean_noise_floor_db"]), rel=1e-5)
        assert row["noise__rain_frame_fraction"] == float(np.mean(g["frame_class"] == 2))
        assert bool(row["rain_actual"]) == (goldens[row["file_key"]][1]["lam"] > 0)
    st_parts = states["rain_detector"].attrs["saved_parquet_files"]
    st_rows = pd.concat([pd.read_parquet(p) for p in st_parts])
    assert sorted(st_rows["file_key"]) == sorted(names)
    for _, st in st_rows.iterrows():
        g = goldens[st["file_key"]][0]
        assert np.array_equal(np.asarray(st["frame_class"], dtype=np.int8), g["frame_class"])
    for _, st in states["noise"].iterrows():       # in-memory remainder: numpy arrays, not lists
        assert np.array_equal(st["is_rain"], goldens[st["file_key"]][0]["frame_class"] == 2)
    assert res.attrs["num_files_processed_total"] == len(names) + 1


def test_td_fast_path_gate_equals_exact(goldens):
    """Default flags decide the TD gate with the float32 filter + exact float64 re-check of the tiles that hold a frame
    within the guard band of td_gate_threshold (2.5e-4 relative + a rounding-error bound that grows with the input's
    low-frequency content).  The gate plane and the labels must equal the float64 path's bit for bit, and the float32
    crest factor must sit far inside the guard band everywhere (asserted: 10x)."""
    from audio_processing_tools_b200.config import build_noise_config
    from audio_processing_tools_b200.engine import BatchEngine
    names = [n for n in goldens if goldens[n][1]["seconds"] == 60][:4]
    clips = [goldens[n][2] for n in names] + [synth_clip_i16(s, 700 + i, lam) for i, (s, lam) in
                                               enumerate(((7.3, 3.0), (33.1, 10.0), (20.6, 0.0), (3.2, 10.0)))]
    n_std = len(clips)
    # hard cases for the float32 filter: rumble 50 dB above the pass-band content (its rounding errors scale with the INPUT),
    # a near-silent clip, and a full-scale clip
    rng = np.random.default_rng(9)
    t = np.arange(int(FS * 30.0)) / FS
    rumble = 0.7 * np.sin(2 * np.pi * 23.0 * t) + 0.002 * rng.standard_normal(t.size)
    clips.append(np.round(np.clip(rumble, -1, 1) * 32767).astype(np.int16))
    clips.append(np.round(rng.standard_normal(int(FS * 22.0)) * 1.5).astype(np.int16))
    clips.append(np.round(np.clip(rng.standard_normal(int(FS * 22.0)) * 0.5, -1, 1) * 32767).astype(np.int16))
    params = default_params(check_duration=3)
    eng = BatchEngine(build_noise_config(FS, params), FS)
    plan, fast = eng.run_clips(clips, ("gate", "td_fast_crest"))
    plan, exact = eng.run_clips(clips, ("gate", "td"))
    assert np.array_equal(fast["gate"], exact["gate"])
    assert np.array_equal(fast["frame_class"], exact["frame_class"])
    assert np.array_equal(fast["event_count"], exact["event_count"])
    assert np.array_equal(fast["clip_stats"], exact["clip_stats"])
    n_std_frames = int(plan.frame_off[n_std])
    c64, c32 = exact["td"][0][:n_std_frames], fast["td_fast_crest"][:n_std_frames]
    ok = (c64 > 0) & (c32 > 0)          # tiles at the clip ends are decided by the exact kernel alone (no float32 value)
    assert ok.mean() > 0.8
    h64, h32 = exact["td"][0][n_std_frames:], fast["td_fast_crest"][n_std_frames:]
    okh = (h64 > 0) & (h32 > 0)
    print(f"hard clips: max relative deviation {float(np.max(np.abs(h32[okh] - h64[okh]) / h64[okh])):.3e}")
    dev = float(np.max(np.abs(c32[ok] - c64[ok]) / c64[ok]))
    near = float(np.mean(np.abs(c64[ok] - 2.5) <= 2.5e-4 * 2.5))
    print(f"float32 crest factor: max relative deviation {dev:.3e} over {int(ok.sum())} frames; {near:.4%} of frames inside the guard band")
    assert dev < 2.5e-5, dev
    for n, c in zip(names, range(len(names))):
        f0, f1 = int(plan.frame_off[c]), int(plan.frame_off[c + 1])
        assert np.array_equal(fast["frame_class"][f0:f1], goldens[n][0]["frame_class"])
    eng.close()


def test_noise_floor_statistics_fallback_paths(oracle_mod):
    """Median / mean of the noise-floor dB values on planes that leave the fast select: a silent clip (every value is eps:
    below the verified range of the monotonicity table, and a constant plane overflows the candidate list), a clip with
    long silent stretches, next to ordinary clips in the same batch.  Exact median, mean within 1e-5 of the oracle."""
    from audio_processing_tools_b200.config import build_noise_config
    from audio_processing_tools_b200.engine import BatchEngine
    rng = np.random.default_rng(5)
    silent = np.zeros(int(FS * 6.0), np.int16)
    gappy = synth_clip_i16(25.0, 811, 3.0)
    gappy[40000:200000] = 0
    tiny = (rng.standard_normal(int(FS * 7.5)) * 1.2).round().astype(np.int16)       # a few LSBs of noise
    clips = [silent, synth_clip_i16(8.0, 812, 10.0), gappy, tiny, synth_clip_i16(41.0, 813, 0.5)]
    params = default_params(check_duration=5)
    eng = BatchEngine(build_noise_config(FS, params), FS)
    plan, out = eng.run_clips(clips, ())
    for c, pcm in enumerate(clips):
        m, s = oracle_mod.run(pcm_to_f32(pcm), dict(params, keep_state_debug=True))
        st = out["clip_stats"][c]
        assert st[7] == np.float32(m["median_noise_floor_db"]), (c, st[7], m["median_noise_floor_db"])
        assert st[6] == pytest.approx(m["mean_noise_floor_db"], rel=1e-5)
        f0, f1 = int(plan.frame_off[c]), int(plan.frame_off[c + 1])
        assert np.array_equal(out["frame_class"][f0:f1], s["frame_class"])
    eng.close()


def test_tensor_core_dft_variant(oracle_mod):
    """DFT-as-GEMM on the tcgen05 tensor cores (fft = "tc": exact int8 limbs of the samples x two fp16 limbs of
    window x twiddle, fp32 accumulation in tensor memory) against the float64 FFT path: band energies within 1e-4
    (north_star tolerance for spectra / band energies in fp32), and -- feeding the full pipeline as the tolerance-path
    front end -- frame labels compared with the oracle (flips counted and bounded, as for the float32 FFT)."""
    from audio_processing_tools_b200.config import build_noise_config
    from audio_processing_tools_b200.engine import BatchEngine
    specs = [(60.0, 900, 3.0), (7.3, 901, 10.0), (33.1, 902, 0.0), (2.5, 903, 3.0), (20.0, 904, 0.5)]
    clips = [synth_clip_i16(s, seed, lam) for s, seed, lam in specs]
    params = default_params(check_duration=2)
    e64 = BatchEngine(build_noise_config(FS, params), FS)
    etc = BatchEngine(build_noise_config(FS, params), FS, fft_f64="tc")
    plan, ref = e64.run_clips(clips, ("band_energy",), full=False)
    plan2, got = etc.run_clips(clips, ("band_energy",), full=False)
    assert etc.L.apt_plan_tc_error(plan2.h) == 0, "a barrier wait of the tensor-core kernel gave up"
    rel = np.abs(got["band_energy"] - ref["band_energy"]) / np.maximum(ref["band_energy"], 1e-12)
    print(f"tensor-core DFT band energies: max relative deviation {float(rel.max()):.3e}")
    assert rel.max() <= 1e-4
    # full pipeline on top of the tensor-core spectra
    plan3, full = etc.run_clips(clips, ())
    assert etc.L.apt_plan_tc_error(plan3.h) == 0
    flips = 0
    for c, pcm in enumerate(clips):
        m, s = oracle_mod.run(pcm_to_f32(pcm), dict(params, keep_state_debug=True))
        f0, f1 = int(plan3.frame_off[c]), int(plan3.frame_off[c + 1])
        flips += int((full["frame_class"][f0:f1] != s["frame_class"]).sum())
        assert full["clip_stats"][c][7] == pytest.approx(m["median_noise_floor_db"], abs=2e-3)
    print(f"tensor-core DFT front end: {flips} of {plan3.nF} frame labels differ from the oracle")
    assert flips <= plan3.nF // 2000
    e64.close(); etc.close()


def test_dump_features_payload_matches_reference():
    """state["features"] under dump_features (rain_signal_processor.py:723-787, :1318): the five per-frame arrays
    decimated by feature_decim, as the unmodified reference returned them (oracle/make_golden_geom.py); with the
    detector debug on, the detector's per-frame features ride along, decimated the same way."""
    import hashlib
    import json
    import os
    from conftest import GOLDEN_DIR
    from audio_processing_tools_b200.edge.rain_signal_processor import RainDetectorProcessor
    g = dict(np.load(os.path.join(GOLDEN_DIR, "features_s61_decim3.npz"), allow_pickle=False))
    meta = json.loads(str(g["meta"]))
    pcm = synth_clip_i16(meta["seconds"], meta["seed"], meta["lam"])
    assert hashlib.sha1(pcm.tobytes()).hexdigest() == meta["pcm_sha1"]
    params = default_params(check_duration=meta["seconds"], dump_features=True, feature_decim=3)
    proc = RainDetectorProcessor()
    _, st = proc.run(pcm_to_f32(pcm), params)
    f = st["features"]
    assert set(f) == {"frame_times", "frame_class", "is_rain", "rain_conf", "noise_conf"}
    for k in f:
        assert f[k].dtype == g["feat_" + k].dtype and np.array_equal(f[k], g["feat_" + k]), k
    # int16 input through the batch entry point gives the same payload; keep_state_features=False drops it
    outs = proc.run_batch([pcm, synth_clip_i16(meta["seconds"] + 1.5, 62, 10.0)], params)
    assert np.array_equal(outs[0][1]["features"]["frame_class"], g["feat_frame_class"])
    assert "features" not in proc.run(pcm, dict(params, keep_state_features=False))[1]
    # with the detector debug on the reference still returns the five arrays: its detector-side dump is an empty dict at
    # feature_dump_level = 0 and takes precedence (rain_signal_processor.py:771-776)
    _, st2 = proc.run(pcm, dict(params, keep_state_debug=True))
    f2 = st2["features"]
    assert set(f2) == set(f) and st2["det_debug"]["feature_dump"] == {}
    assert np.array_equal(f2["frame_class"], g["feat_frame_class"])


def test_det_debug_key_set_and_soft_labels_match_reference():
    """state["det_debug"] carries every key the reference's detector exports under default dump flags
    (rain_frame_classifier.py:1000-1047), the scalar echoes have the reference's values, and the soft TD label arrays
    under td_soft_enable (:85-110) are equal (tests/golden/detdebug_s63.npz from oracle/make_golden_geom.py)."""
    import hashlib
    import json
    import os
    from conftest import GOLDEN_DIR
    from audio_processing_tools_b200.edge.rain_signal_processor import RainDetectorProcessor
    g = dict(np.load(os.path.join(GOLDEN_DIR, "detdebug_s63.npz"), allow_pickle=False))
    meta = json.loads(str(g["meta"]))
    pcm = synth_clip_i16(meta["seconds"], meta["seed"], meta["lam"])
    assert hashlib.sha1(pcm.tobytes()).hexdigest() == meta["pcm_sha1"]
    params = default_params(check_duration=meta["seconds"], keep_state_debug=True)
    params["detector"] = dict(params["detector"], td_soft_enable=True, td_soft_crest_factor_min=3.0, td_soft_kurtosis_min=4.0)
    _, st = RainDetectorProcessor().run(pcm_to_f32(pcm), params)
    dd = st["det_debug"]
    assert set(dd) == set(meta["keys"]), sorted(set(dd) ^ set(meta["keys"]))
    for k, v in meta["scalars"].items():
        got = dd[k]
        got = got.item() if isinstance(got, (np.floating, np.integer, np.bool_)) else got
        assert (str(got) == v) if isinstance(v, str) and not isinstance(got, str) else (got == v), (k, got, v)
    for k in ("td_vote_count", "td_soft_score", "td_soft_label", "raw_spectral_dump_mask", "sparse_frame_idx"):
        assert dd[k].dtype == g[k].dtype and np.array_equal(dd[k], g[k]), k
    assert int(g["td_vote_count"].sum()) > 0
    # the suppressor-side debug dictionary and the state itself: same keys, except the per-frame statistics of the gain
    # stages (debug["gain_dbg"]: medians / percentiles of G_raw, G_freq, G_time), which are not produced
    assert set(meta["debug_keys"]) - set(st["debug"]) == {"gain_dbg"}
    assert set(st["debug"]) <= set(meta["debug_keys"])
    assert set(st) == set(meta["state_keys"])


def test_feature_dump_payload_matches_reference():
    """dump_features with feature_dump_level = 1 (dense + sparse detector dump, soft TD votes), feature_decim = 2: the
    payload's key set equals the reference's and every array matches (integer / boolean arrays and the TD features bit for
    bit, kurtosis and the raw spectral features at their tolerance) -- tests/golden/featuredump_s64.npz."""
    import hashlib
    import json
    import os
    from conftest import GOLDEN_DIR
    from audio_processing_tools_b200.edge.rain_signal_processor import RainDetectorProcessor
    g = dict(np.load(os.path.join(GOLDEN_DIR, "featuredump_s64.npz"), allow_pickle=False))
    meta = json.loads(str(g["meta"]))
    pcm = synth_clip_i16(meta["seconds"], meta["seed"], meta["lam"])
    assert hashlib.sha1(pcm.tobytes()).hexdigest() == meta["pcm_sha1"]
    params = default_params(check_duration=meta["seconds"], dump_features=True, feature_decim=meta["feature_decim"])
    params["detector"] = dict(params["detector"], **meta["detector"])
    _, st = RainDetectorProcessor().run(pcm_to_f32(pcm), params)
    f = st["features"]
    assert "det_debug" not in st                      # the dump rides in the features payload, not in the debug state
    assert {"feat_" + k for k in f} == set(g) - {"meta"}
    for k, v in f.items():
        ref = g["feat_" + k]
        assert v.shape == ref.shape and v.dtype == ref.dtype, k
        if v.dtype.kind in "biu" or k.startswith("td_block") or k in ("td_crest_factor", "frame_times", "rain_conf", "noise_conf"):
            assert np.array_equal(v, ref), k
        elif k == "td_kurtosis":
            np.testing.assert_allclose(v, ref, rtol=2e-6, atol=1e-6, err_msg=k)
        elif k.startswith("sparse_raw"):
            np.testing.assert_allclose(v, ref, rtol=1e-6, atol=1e-30, err_msg=k)
        else:
            np.testing.assert_allclose(v, ref, rtol=1e-6, atol=1e-7, err_msg=k)
    assert f["sparse_frame_idx"].size > 0
