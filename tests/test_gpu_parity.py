"""GPU parity: the CUDA path (through the C ABI) against the golden vectors produced by the
reference and against the CPU oracle on the same seeded inputs.

Bars (north_star): frame classes, event indices and counts bit-exact; float planes within the
stated tolerance (<= 1e-4 relative; spectra relative to the frame maximum).  With the default
float64 FFT the integer outputs AND the float32 planes that feed decisions are expected to be
bit-equal to the oracle; the asserted float tolerances are the ones written next to each check.
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from audio_processing_tools_b200.synth import default_params, pcm_to_f32, synth_clip_i16

pytestmark = pytest.mark.gpu

ALL_PLANES = ("S", "P", "det_noise_psd", "det_noise_lag", "D", "noise_psd", "mode_flux", "norm_flux",
              "score", "td", "raw", "band_energy", "gate", "x_td", "G", "ratio_med", "S_hat")


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch


def make_engine(params, **kw):
    from audio_processing_tools_b200.config import build_noise_config
    from audio_processing_tools_b200.engine import BatchEngine
    cfg = build_noise_config(int(params.get("sample_rate", 11162)), params)
    return BatchEngine(cfg, int(params.get("sample_rate", 11162)),
                       clip_rain_min_frames=int(params.get("clip_rain_min_frames", 1)), **kw)


def frame_rel_err(a, b):
    """max |a-b| per frame relative to the frame's max |b| (spectra tolerance of SURVEY 7.2)."""
    num = np.abs(a - b).reshape(a.shape[0], -1).max(axis=1)
    den = np.abs(b).reshape(b.shape[0], -1).max(axis=1) + 1e-30
    return float((num / den).max())


@pytest.mark.parametrize("name", golden_names(min_level=2))
def test_full_planes_match_reference_golden(torch_cuda, name):
    g, meta, pcm, params = load_golden(name)
    eng = make_engine(params)
    snr = "snr_gate_t" in g
    plan, out = eng.run_clips([pcm], ALL_PLANES + (("snr_mode", "snr_gate") if snr else ()))
    T = plan.nF
    if snr:     # frame SNR over the mode bins and its gate (rain_signal_processor.py:1064-1077)
        np.testing.assert_allclose(out["snr_mode"], g["snr_mode_t"], rtol=1e-5)
        np.testing.assert_allclose(out["snr_gate"], g["snr_gate_t"], rtol=1e-5, atol=1e-7)
    S = out["S"].view(np.complex64).reshape(T, -1)
    # spectra: <= 1e-6 of the frame maximum (float64 FFT rounded to complex64; bit-equal in practice)
    assert frame_rel_err(S, g["S"]) <= 1e-6
    mism = int((S != g["S"]).sum())
    assert mism <= max(2, S.size // 100000), f"{mism} complex64 bins differ from the reference"
    # noise PSD planes: <= 1e-4 relative (bit-equal when S is)
    np.testing.assert_allclose(out["det_noise_psd"], g["detector_noise_psd_band"], rtol=1e-4, atol=1e-12)
    np.testing.assert_allclose(out["det_noise_lag"], g["detector_noise_psd_lag_band"], rtol=1e-4, atol=1e-12)
    np.testing.assert_allclose(out["noise_psd"], g["noise_psd_band"], rtol=1e-4, atol=1e-12)
    # labels / events: bit-exact
    assert np.array_equal(out["frame_class"], g["frame_class"])
    n = int(out["event_count"][0])
    assert np.array_equal(out["event_idx"][:n], g["event_idx"])
    assert np.array_equal(out["rain_conf"], g["rain_conf"])
    assert np.array_equal(out["noise_conf"], g["noise_conf"])
    # detector features: <= 1e-4 relative
    for i, k in enumerate(("primary_mode_flux", "support_mode_flux_1", "support_mode_flux_2",
                           "support_mode_flux_3", "support_mode_flux_4")):
        np.testing.assert_allclose(out["norm_flux"][i], g["det_" + k], rtol=1e-4, atol=1e-5, err_msg=k)
    np.testing.assert_allclose(out["score"], g["det_mode_flux_score"], rtol=1e-4, atol=1e-5)
    assert np.array_equal(out["gate"].astype(bool), g["det_td_gate_mask"])
    for i, k in enumerate(("td_crest_factor", "td_kurtosis", "td_block_energy_crest",
                           "td_block_peak_width_50", "td_block_post_pre_energy_ratio")):
        np.testing.assert_allclose(out["td"][i], g["det_" + k], rtol=1e-4, atol=1e-5, err_msg=k)
    from audio_processing_tools_b200.edge.rain_signal_processor import RAW_SPECTRAL_FEATURE_NAMES
    for i, k in enumerate(RAW_SPECTRAL_FEATURE_NAMES):
        np.testing.assert_allclose(out["raw"][i], g["det_" + k], rtol=1e-4, atol=1e-6, err_msg=k)
    st = out["clip_stats"][0]
    assert int(st[1]) == int(g["metric_rain_frame_count"])
    assert st[6] == pytest.approx(float(g["metric_mean_noise_floor_db"]), rel=1e-5)
    assert st[7] == pytest.approx(float(g["metric_median_noise_floor_db"]), rel=1e-6)
    # suppressor gain (rain_signal_processor.py:400-533): <= 1e-5 (float32 sums of a 3-tap kernel in another order)
    np.testing.assert_allclose(out["G"], g["G_band"], rtol=1e-5, atol=2e-7)
    np.testing.assert_allclose(out["ratio_med"], g["np_ratio_median_t"], rtol=1e-6, atol=1e-12)
    G_full = np.ones((T, S.shape[1]), np.float32)
    G_full[:, 10:81] = g["G_band"]
    S_hat = out["S_hat"].view(np.complex64).reshape(T, -1)
    np.testing.assert_allclose(S_hat, G_full * g["S"], rtol=1e-5, atol=1e-9)
    eng.close()


@pytest.mark.parametrize("name", [n for n in golden_names() if n not in golden_names(min_level=2)])
def test_events_match_reference_golden(torch_cuda, name):
    """Through the reference-facing processor API (RainDetectorProcessor.run), int16-derived float32 input."""
    from audio_processing_tools_b200.edge.rain_signal_processor import RainDetectorProcessor
    g, meta, pcm, params = load_golden(name)
    params["keep_state_debug"] = True
    proc = RainDetectorProcessor()
    m, s = proc.run(pcm_to_f32(pcm), params)
    assert np.array_equal(s["frame_class"], g["frame_class"])
    assert np.array_equal(np.flatnonzero(s["frame_class"] == 2).astype(np.int32), g["event_idx"])
    assert np.array_equal(s["rain_conf"], g["rain_conf"])
    assert np.array_equal(s["noise_conf"], g["noise_conf"])
    assert np.array_equal(s["times"], g["times"])
    for k in ("rain_frame_count", "clip_is_rain", "clip_rain_conf", "median_rain_conf",
              "clip_rain_fraction", "clip_rain_min_frames"):
        assert m[k] == g["metric_" + k].item(), k
    assert m["mean_noise_floor_db"] == pytest.approx(float(g["metric_mean_noise_floor_db"]), rel=1e-5)
    assert m["median_noise_floor_db"] == pytest.approx(float(g["metric_median_noise_floor_db"]), rel=1e-6)
    if "det_primary_mode_flux" in g:
        dd = s["det_debug"]
        for k in ("primary_mode_flux", "support_mode_flux_1", "support_mode_flux_2", "support_mode_flux_3",
                  "mode_flux_score", "td_crest_factor", "td_kurtosis", "td_block_energy_crest",
                  "td_block_peak_width_50", "td_block_post_pre_energy_ratio"):
            np.testing.assert_allclose(dd[k], g["det_" + k], rtol=1e-4, atol=1e-5, err_msg=k)
        assert np.array_equal(dd["td_gate_mask"], g["det_td_gate_mask"])


def test_ragged_batch_matches_oracle_i16(torch_cuda, oracle_mod):
    """Ragged multi-clip batch, int16 wire input, every plane against the CPU oracle."""
    specs = [(7.3, 201, 3.0), (5.0, 202, 0.0), (11.9, 203, 10.0), (6.02, 204, 0.5), (9.5, 205, 10.0)]
    clips = [synth_clip_i16(s, seed, lam) for s, seed, lam in specs]
    params = default_params(check_duration=5)
    eng = make_engine(params)
    plan, out = eng.run_clips(clips, ALL_PLANES)
    flips = 0
    for c, pcm in enumerate(clips):
        p2 = dict(params, keep_state_debug=True)
        m, s = oracle_mod.run(pcm_to_f32(pcm), p2)
        f0, f1 = int(plan.frame_off[c]), int(plan.frame_off[c + 1])
        s0, s1 = int(plan.sample_off[c]), int(plan.sample_off[c + 1])
        T = f1 - f0
        assert T == s["frame_class"].size
        flips += int((out["frame_class"][f0:f1] != s["frame_class"]).sum())
        n = int(out["event_count"][c])
        assert np.array_equal(out["event_idx"][f0:f0 + n], s["event_idx"])
        S = out["S"][f0:f1].view(np.complex64).reshape(T, -1)
        assert frame_rel_err(S, s["S"]) <= 1e-6
        np.testing.assert_allclose(out["P"][f0:f1][:, 10:81], s["P_band"], rtol=1e-4, atol=1e-12)
        np.testing.assert_allclose(out["D"][f0:f1], s["D_band"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(out["det_noise_psd"][f0:f1], s["N1_band"], rtol=1e-4, atol=1e-12)
        np.testing.assert_allclose(out["noise_psd"][f0:f1], s["N2_band"], rtol=1e-4, atol=1e-12)
        np.testing.assert_allclose(out["mode_flux"][:, f0:f1], s["mode_flux"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(out["norm_flux"][:, f0:f1], s["norm_flux"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(out["x_td"][s0:s1], s["x_td"], rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(out["td"][:, f0:f1], s["td"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(out["raw"][:, f0:f1], s["raw"], rtol=1e-4, atol=1e-6)
        st = out["clip_stats"][c]
        assert int(st[1]) == m["rain_frame_count"]
        assert st[6] == pytest.approx(m["mean_noise_floor_db"], rel=1e-5)
        assert st[7] == pytest.approx(m["median_noise_floor_db"], rel=1e-6)
    assert flips == 0, f"{flips} frame labels differ from the oracle"
    eng.close()


def test_default_planes_mode_lane_path(torch_cuda, oracle_mod):
    """Default flag set (no pass-1 debug planes): pass 1 tracks the mode bins only.  Labels, events and the
    flux planes must equal the oracle bit for bit; ragged lengths exercise the serial kernels' tails."""
    specs = [(20.0, 300, 0.0), (13.37, 301, 3.0), (20.0, 302, 10.0), (5.5, 303, 10.0), (9.0, 304, 0.5),
             (20.0, 305, 3.0), (6.1, 306, 3.0)]
    clips = [synth_clip_i16(s, seed, lam) for s, seed, lam in specs]
    params = default_params(check_duration=5)
    eng = make_engine(params)
    plan, out = eng.run_clips(clips, ("mode_flux", "norm_flux", "score", "noise_psd"))
    for c, pcm in enumerate(clips):
        m, s = oracle_mod.run(pcm_to_f32(pcm), dict(params, keep_state_debug=True))
        f0, f1 = int(plan.frame_off[c]), int(plan.frame_off[c + 1])
        assert np.array_equal(out["frame_class"][f0:f1], s["frame_class"])
        assert np.array_equal(out["rain_conf"][f0:f1], s["rain_conf"])
        assert np.array_equal(out["noise_conf"][f0:f1], s["noise_conf"])
        n = int(out["event_count"][c])
        assert n == m["rain_frame_count"] and np.array_equal(out["event_idx"][f0:f0 + n], s["event_idx"])
        assert np.array_equal(out["mode_flux"][:, f0:f1], s["mode_flux"])
        assert np.array_equal(out["norm_flux"][:, f0:f1], s["norm_flux"])
        assert np.array_equal(out["score"][f0:f1], s["score"])
        assert np.array_equal(out["noise_psd"][f0:f1], s["N2_band"])
        st = out["clip_stats"][c]
        assert st[6] == pytest.approx(m["mean_noise_floor_db"], rel=1e-6)
        assert st[7] == m["median_noise_floor_db"]
    # the same batch without any optional plane (dB plane produced in place) gives the same statistics
    plan2, out2 = eng.run_clips(clips, ())
    assert np.array_equal(out2["frame_class"], out["frame_class"])
    assert np.array_equal(out2["clip_stats"], out["clip_stats"])
    assert np.array_equal(out2["event_count"], out["event_count"])
    eng.close()


def test_f32_fft_mode_events(torch_cuda, oracle_mod):
    """float32 FFT variant (north_star subsystem 2): spectra within 1e-6 of frame max, labels equal."""
    clips = [synth_clip_i16(20, 300 + i, (0.0, 3.0, 10.0)[i]) for i in range(3)]
    params = default_params(check_duration=20)
    eng = make_engine(params, fft_f64=False)
    plan, out = eng.run_clips(clips, ("S",))
    flips = 0
    for c, pcm in enumerate(clips):
        m, s = oracle_mod.run(pcm_to_f32(pcm), dict(params, keep_state_debug=True))
        f0, f1 = int(plan.frame_off[c]), int(plan.frame_off[c + 1])
        S = out["S"][f0:f1].view(np.complex64).reshape(f1 - f0, -1)
        assert frame_rel_err(S, s["S"]) <= 2e-6
        flips += int((out["frame_class"][f0:f1] != s["frame_class"]).sum())
    assert flips == 0
    eng.close()


def test_features_only_stage(torch_cuda, oracle_mod):
    """BASELINE config 2 path: STFT + band energies only (no detector buffers needed)."""
    pcm = synth_clip_i16(30, 400, 3.0)
    params = default_params(check_duration=30)
    eng = make_engine(params)
    plan, out = eng.run_clips([pcm], ("band_energy", "P"), full=False)
    s = oracle_mod.process(pcm_to_f32(pcm), dict(params))
    P = out["P"]
    np.testing.assert_allclose(P[:, 10:81], s["P_band"], rtol=1e-4, atol=1e-12)
    be = out["band_energy"]
    P64 = P.astype(np.float64)
    for i, (lo, hi) in enumerate(((11, 14), (19, 24), (35, 41), (54, 58), (73, 76))):
        np.testing.assert_allclose(be[i], P64[:, lo:hi + 1].sum(axis=1), rtol=1e-6)
    np.testing.assert_allclose(be[5], P64[:, 10:81].sum(axis=1) + 1e-9, rtol=1e-6)
    eng.close()


def test_host_path_and_errors(torch_cuda):
    """apt_run_host_i16 (e2e path) equals the device path; bad inputs fail loudly."""
    from audio_processing_tools_b200.engine import AptError
    from audio_processing_tools_b200.edge.rain_signal_processor import RainDetectorProcessor
    clips = [synth_clip_i16(6 + i, 500 + i, 3.0) for i in range(10)]
    params = default_params(check_duration=6)
    eng = make_engine(params)
    plan, out = eng.run_clips(clips, ())
    cat = np.concatenate(clips)
    host = {"frame_class": np.zeros(plan.nF, np.int8), "rain_conf": np.zeros(plan.nF, np.float32),
            "noise_conf": np.zeros(plan.nF, np.float32), "event_idx": np.zeros(plan.nF, np.int32),
            "event_count": np.zeros(plan.n_clips, np.int32), "clip_stats": np.zeros((plan.n_clips, 8), np.float32)}
    eng.run_host_i16(plan, cat, host)
    assert np.array_equal(host["frame_class"], out["frame_class"])
    assert np.array_equal(host["event_count"], out["event_count"])
    assert np.array_equal(host["clip_stats"], out["clip_stats"])
    with pytest.raises(AptError):
        eng.plan_for([100])          # shorter than one frame
    eng.close()
    proc = RainDetectorProcessor()
    with pytest.raises(TypeError):
        proc.run([0.0] * 1000, params)
    with pytest.raises(ValueError):
        proc.run(np.zeros(1000, np.float32), params)
    with pytest.raises(AttributeError):
        proc.run(np.zeros(6 * 11162, np.float32), {"sample_rate": 11162, "check_duration": 6})
    with pytest.raises(NotImplementedError):
        proc.run(np.zeros(6 * 11162, np.float32), dict(params, process_dtype="float64"))


@pytest.mark.parametrize("n_fft,hop", [(256, 64), (512, 256), (1024, 256), (2048, 1024), (4096, 1024)])
@pytest.mark.parametrize("fft_f64", [True, False])
def test_frame_size_sweep_features(torch_cuda, n_fft, hop, fft_f64):
    """BASELINE config 5 (features stage): generic power-of-two STFT against the reference's spectra and
    band-energy features; the float32 FFT variant within 1e-5 of the frame maximum."""
    from test_oracle_golden import load_sweep
    from audio_processing_tools_b200.edge.rain_signal_processor import RAW_SPECTRAL_FEATURE_NAMES
    from audio_processing_tools_b200.engine import AptError
    g, meta, params = load_sweep(n_fft, hop)
    eng = make_engine(params, fft_f64=fft_f64)
    plan, out = eng.run_clips([g["pcm"]], ("S", "P", "raw", "band_energy"), full=False)
    T, F = g["S"].shape
    assert plan.nF == T
    S = out["S"].view(np.complex64).reshape(T, F)
    assert frame_rel_err(S, g["S"]) <= (1e-6 if fft_f64 else 1e-5)
    if fft_f64:
        assert (S != g["S"]).mean() < 2e-3
    P_ref = (np.abs(g["S"]).astype(np.float32)) ** 2
    np.testing.assert_allclose(out["P"], P_ref, rtol=1e-4, atol=1e-6 * float(P_ref.max()))
    if fft_f64:
        for i, k in enumerate(RAW_SPECTRAL_FEATURE_NAMES):
            np.testing.assert_allclose(out["raw"][i], g["det_" + k], rtol=1e-4, atol=1e-6, err_msg=k)
    band = g["band_mask"]
    np.testing.assert_allclose(out["band_energy"][-1], out["P"][:, band].astype(np.float64).sum(axis=1) + 1e-9, rtol=1e-6)
    eng.close()
    if (n_fft, hop) == (256, 64):          # the full pipeline needs a hop that is a multiple of 64
        eng = make_engine(dict(params, hop=96), fft_f64=fft_f64)
        with pytest.raises(AptError):
            eng.run_clips([g["pcm"]], ())
        eng.close()


@pytest.mark.parametrize("n_fft,hop", [(512, 256), (1024, 256), (2048, 1024), (4096, 1024), (256, 256), (512, 128), (256, 64), (512, 192)])
def test_full_pipeline_other_frame_sizes(torch_cuda, n_fft, hop):
    """The whole detector at frame sizes other than 256 / 128 against the unmodified reference (geom_* fixtures): the
    TD crest factor and gate bit for bit; labels, confidences and events identical; per-mode flux, score and the
    noise-floor statistics within the float32 spectrum tolerance (the generic FFT is another algorithm than pocketfft,
    so single power values differ in the last bit).  Batched with a second, ragged clip to cover the offsets."""
    from conftest import load_golden
    from audio_processing_tools_b200.synth import synth_clip_i16
    g, meta, pcm, params = load_golden(f"geom_nfft{n_fft}_hop{hop}")
    other = synth_clip_i16(7.3, 900 + n_fft // 256, 10.0)
    eng = make_engine(params)
    plan, out = eng.run_clips([other, pcm], ("td", "gate", "score", "norm_flux"))
    f0, f1 = int(plan.frame_off[1]), int(plan.frame_off[2])
    T = g["frame_class"].size
    assert f1 - f0 == T
    assert np.array_equal(out["td"][0][f0:f1], g["det_td_crest_factor"])
    assert np.all(np.isnan(out["td"][1:]))
    assert np.array_equal(out["gate"][f0:f1].astype(bool), g["det_td_gate_mask"])
    np.testing.assert_allclose(out["score"][f0:f1], g["det_mode_flux_score"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["norm_flux"][0][f0:f1], g["det_primary_mode_flux"], rtol=1e-4, atol=1e-5)
    assert np.array_equal(out["frame_class"][f0:f1], g["frame_class"])
    assert np.array_equal(out["rain_conf"][f0:f1], g["rain_conf"])
    assert np.array_equal(out["noise_conf"][f0:f1], g["noise_conf"])
    n_ev = int(out["event_count"][1])
    assert np.array_equal(out["event_idx"][f0:f0 + n_ev], g["event_idx"])
    st = out["clip_stats"][1]
    assert st[6] == pytest.approx(g["metric_mean_noise_floor_db"].item(), rel=1e-5)
    assert st[7] == pytest.approx(g["metric_median_noise_floor_db"].item(), rel=1e-5)
    # the host path (pageable clips in, labels out) gives the same labels
    _, res = eng.run_host_clips([other, pcm])
    assert np.array_equal(res["frame_class"][f0:f1], g["frame_class"])
    eng.close()


def test_exact_div_sqrt(torch_cuda):
    """The branch-free float32 division / square root used in the power and normalisation paths are
    bit-identical to the IEEE intrinsics: exhaustive for sqrt on [1, 2], 2^30 hashed pairs for division."""
    import ctypes as C
    from audio_processing_tools_b200 import _lib
    L = _lib.load()
    ctx = C.c_void_p()
    assert L.apt_init(0, C.byref(ctx)) == 0
    bad = C.c_int64(-1)
    assert L.apt_selftest(ctx, 0, 0, C.byref(bad)) == 0 and bad.value == 0
    bad = C.c_int64(-1)
    assert L.apt_selftest(ctx, 1, 1 << 30, C.byref(bad)) == 0 and bad.value == 0
    L.apt_destroy(ctx)


def test_dsd_emulator_matches_reference(torch_cuda):
    """the GPU drop-size-distribution emulator against the unmodified reference's per-minute
    100-vectors (integer histograms: bit-exact), one batch holding every case with the same windowing."""
    from test_dsd_oracle import case_pcm, dsd_cases
    from audio_processing_tools_b200.host_analysis.device_dsd_processing_emulator import DsdProcessingEmualtor
    cases = dsd_cases()
    for win in (False, True):
        sel = [(m, ref) for m, ref in cases if m["window"] == win]
        em = DsdProcessingEmualtor(fs=11162, frame_length=512, hop_length=512, bwindow=win)
        outs = em.process_audio_batch([case_pcm(m) for m, _ in sel], [m["ts"] for m, _ in sel])
        for (m, ref), out in zip(sel, outs):
            assert len(out) == ref.shape[0], m["name"]
            assert np.array_equal(np.asarray(out).reshape(ref.shape), ref), m["name"]
    # float input scaled as parse.pcm_to_float is accepted, anything else is refused
    m, ref = cases[0]
    pcm = case_pcm(m)
    one = DsdProcessingEmualtor().process_audio_data(pcm.astype(np.float64) / 32768.0, m["ts"])
    assert np.array_equal(np.asarray(one), ref)
    with pytest.raises(ValueError):
        DsdProcessingEmualtor().process_audio_data(pcm.astype(np.float64) / 32767.0, 0)


def test_dsd_fft_kernels_agree(torch_cuda, monkeypatch):
    """The 16-lanes-per-frame FFT and the warp-per-clip state machine of the drop-size emulator against the generic
    one-CTA-per-frame FFT and the thread-per-clip machine (APT_DSD_FFT_GENERIC=1, APT_DSD_STATE_SERIAL=1) on clips with
    odd lengths and timestamps inside a minute: the same per-minute vectors."""
    from audio_processing_tools_b200.host_analysis.device_dsd_processing_emulator import DsdProcessingEmualtor
    clips = [synth_clip_i16(130.3, 201, 10.0), synth_clip_i16(59.0, 202, 30.0), synth_clip_i16(0.03, 203, 3.0),
             synth_clip_i16(75.7, 204, 0.0), synth_clip_i16(190.0, 205, 3.0)]
    ts = [0.0, 17.25, 3.0, 59.99, 1714564800.5]
    for win in (False, True):
        em = DsdProcessingEmualtor(fs=11162, frame_length=512, hop_length=512, bwindow=win)
        monkeypatch.setenv("APT_DSD_FFT_GENERIC", "1")
        monkeypatch.setenv("APT_DSD_STATE_SERIAL", "1")
        old = em.process_audio_batch(clips, ts)
        monkeypatch.delenv("APT_DSD_FFT_GENERIC")
        monkeypatch.delenv("APT_DSD_STATE_SERIAL")
        new = em.process_audio_batch(clips, ts)
        assert [len(o) for o in old] == [len(o) for o in new]
        for o, n in zip(old, new):
            assert np.array_equal(np.asarray(o), np.asarray(n))


def test_transform_dsd_entry_points(torch_cuda):
    import datetime as dt
    from test_dsd_oracle import case_pcm, dsd_cases
    from audio_processing_tools_b200 import transform
    m, ref = dsd_cases()[3]          # 150 s clip: transform processes its first 60 s
    pcm = case_pcm(m)
    meta = {"sample_rate": 11162, "device_id": "C00001", "time": dt.datetime(2024, 5, 1, 12, 0, 0)}
    df = transform.process_audio_file_dsd("k0", pcm, meta)
    assert len(df) == 1 and df["device"].iloc[0] == "C00001" and df["key"].iloc[0] == "k0"
    assert np.array_equal(df[[f"dsd{i}" for i in range(32)]].to_numpy()[0], ref[0, :32])
    assert np.array_equal(df[[f"pft{i}" for i in range(30)]].to_numpy()[0], ref[0, 32:62])
    assert np.array_equal(df[[f"fft{i}" for i in range(38)]].to_numpy()[0], ref[0, 62:])
    assert df["weighted_dsd_sum"].iloc[0] == pytest.approx(float((ref[0, :32] * np.array(list(transform.dsd_weights.values()))).sum()))
    assert df["time"].iloc[0] == meta["time"] + dt.timedelta(minutes=1)


def test_mark3_loader_feeds_host_path(torch_cuda):
    """Mark-3 files -> pinned int16 batch -> apt_run_host_i16, equal to the device-resident path."""
    from audio_processing_tools_b200 import parse
    clips = [synth_clip_i16(6.0 + 0.5 * i, 700 + i, (0.5, 3.0, 10.0)[i % 3]) for i in range(6)]
    files = [parse.build_mark_audio_file(c, ts=1700000000 + i, device_id=f"C{i:06d}") for i, c in enumerate(clips)]
    loader = parse.Mark3BatchLoader(sum(c.size for c in clips))
    pcm, lengths, metas = loader.load(files)
    assert loader.pinned and [m["device_id"] for m in metas] == [f"C{i:06d}" for i in range(6)]
    params = default_params(check_duration=6)
    eng = make_engine(params)
    plan = eng.plan_for(list(lengths))
    host = {"frame_class": np.zeros(plan.nF, np.int8), "rain_conf": None, "noise_conf": None,
            "event_idx": np.zeros(plan.nF, np.int32), "event_count": np.zeros(plan.n_clips, np.int32),
            "clip_stats": np.zeros((plan.n_clips, 8), np.float32)}
    eng.run_host_i16(plan, pcm, host)
    plan2, out = eng.run_clips(clips, ())
    assert np.array_equal(host["frame_class"], out["frame_class"])
    assert np.array_equal(host["event_count"], out["event_count"])
    assert np.array_equal(host["clip_stats"], out["clip_stats"])
    eng.close()


@pytest.mark.parametrize("name,seconds,seed,lam", [("audio_s14_l3_6s", 6.0, 14, 3.0), ("audio_s15_l10_5s", 5.03, 15, 10.0)])
def test_output_audio_matches_reference(torch_cuda, name, seconds, seed, lam):
    """compute_output_audio: gain -> S_hat -> inverse STFT (rain_signal_processor.py:1113-1128) through the
    processor API, against the reference run (its ISTFT is the harness' librosa stand-in: unpinned link).
    Tolerance: 2e-6 absolute on a waveform of amplitude <= 1 (float32 output of float64 overlap-add)."""
    import os
    from conftest import GOLDEN_DIR
    from audio_processing_tools_b200.edge.rain_signal_processor import RainDetectorProcessor
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    pcm = synth_clip_i16(seconds, seed, lam)
    params = default_params(check_duration=int(seconds), keep_state_audio=True, keep_state_spectra=True, keep_state_debug=True)
    m, s = RainDetectorProcessor().run(pcm_to_f32(pcm), params)
    assert np.array_equal(s["frame_class"], g["frame_class"])
    assert s["output_audio"].dtype == np.float32 and s["output_audio"].shape == g["output_audio"].shape
    np.testing.assert_allclose(s["filtered_audio"], g["filtered_audio"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(s["output_audio"], g["output_audio"], rtol=1e-5, atol=2e-6)
    # (not bit-equal: scipy's irfft runs in float32 on complex64 input, the kernel inverts in float64)
    assert float(np.abs(s["output_audio"] - g["output_audio"]).max()) < 2e-6


@pytest.mark.parametrize("name", ["peaks_s16_l10_20s", "peaks_s17_l3_15s_alt"])
def test_peak_features_match_reference(torch_cuda, name):
    """Optional peak-structure features (rain_frame_classifier.py:761-843, peak_features_enable): integer counts and
    the binary gate bit-exact, the top-P ratio equal as float32, against the unmodified reference."""
    import json
    import os
    from conftest import GOLDEN_DIR
    from audio_processing_tools_b200.edge.rain_signal_processor import RainDetectorProcessor
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(g["meta"]))
    pcm = synth_clip_i16(meta["seconds"], meta["seed"], meta["lam"])
    params = default_params(check_duration=int(meta["seconds"]), keep_state_debug=True)
    params["detector"].update({"peak_features_enable": True, **meta["detector_extra"]})
    m, s = RainDetectorProcessor().run(pcm_to_f32(pcm), params)
    dd = s["det_debug"]
    assert np.array_equal(s["frame_class"], g["frame_class"])
    assert dd["peak_features_enable"] is True
    assert np.array_equal(dd["peak_valid_count"], g["peak_valid_count"])
    assert np.array_equal(dd["peak_count_by_mode"], g["peak_count_by_mode"])
    assert np.array_equal(dd["peak_gate_score"], g["peak_gate_score"])
    assert np.array_equal(dd["peak_ratio"], g["peak_ratio"])


@pytest.mark.parametrize("idx", range(6))
def test_band_noise_estimator_matches_reference(torch_cuda, idx):
    """the GPU band noise estimator through BandNoiseEstimatorProcessor.run against the unmodified
    reference: rain masks / FFT rain flags / counters bit-exact, float64 series within 1e-9 relative
    (block-wise streaming filters and another FFT algorithm: 1e-13-level differences before accumulation)."""
    from test_band_noise_oracle import band_cases, case_pcm, check_against_golden
    from audio_processing_tools_b200.edge.band_noise_processor import BandNoiseEstimatorProcessor
    g, meta = band_cases()
    m = meta[idx]
    params = {"sample_rate": 11162, "check_duration": m["seconds"], **m["extra"]}
    res, st = BandNoiseEstimatorProcessor().run(pcm_to_f32(case_pcm(m)), params)
    check_against_golden(st, g, m["name"], rtol=1e-9)
    import json
    ref = json.loads(str(g[m["name"] + "__results"]))
    for k, v in ref.items():
        if isinstance(v, str):
            assert res[k] == v, k
        elif isinstance(v, int):
            assert res[k] == v, k
        else:
            assert res[k] == pytest.approx(v, rel=1e-9, nan_ok=True), k


def test_band_noise_batch_and_int16(torch_cuda):
    """A ragged batch (int16 wire input) equals the per-clip runs; clips shorter than one frame give empty outputs."""
    from audio_processing_tools_b200.edge.band_noise_processor import BandNoiseEstimatorProcessor
    clips = [synth_clip_i16(20.0, 81, 3.0), synth_clip_i16(0.03, 82, 3.0), synth_clip_i16(31.4, 83, 10.0)]
    proc = BandNoiseEstimatorProcessor()
    params = {"sample_rate": 11162}
    batch = proc.run_batch(clips, params)
    assert batch[1][0]["n_frames"] == 0 and np.isnan(batch[1][0]["gain_med"])
    for c, (res, st) in zip(clips, batch):
        r1, s1 = proc.run(pcm_to_f32(c), params)
        assert r1["n_frames"] == res["n_frames"]
        for k in ("M_clean", "N_E", "G_mag", "subE"):
            assert np.array_equal(s1[k], st[k]), k
        assert np.array_equal(s1["rain_submask"], st["rain_submask"])


def test_band_noise_warp_kernel_equals_serial(torch_cuda, monkeypatch):
    """The band noise estimator's wavefront filters (one lane per second-order section), warp-per-clip state machine (ring
    buffer and its sorted copy spread over the lanes) and 16- / 8-lanes-per-frame FFT against the kernels they replace
    (APT_BNE_FILTER_SERIAL / APT_BNE_STATE_SERIAL / APT_BNE_FFT_GENERIC keep those): with the same FFT every output is
    bit-equal -- all arrays, the counters and the adaptive quantile; the two FFT algorithms agree to 1e-9 relative on
    the float64 series and give the same masks and flags."""
    from audio_processing_tools_b200.edge.band_noise_processor import BandNoiseEstimatorProcessor
    clips = [synth_clip_i16(41.0, 181, 3.0), synth_clip_i16(0.03, 182, 3.0), synth_clip_i16(63.7, 183, 30.0),
             synth_clip_i16(25.2, 184, 0.0), synth_clip_i16(12.0, 185, 10.0)]
    variants = [{},                                                                   # frame 512, S = 4
                {"noise_buffer_ttl_frames": 12, "W": 24, "W_min": 6},                 # entries expire
                {"W": 64, "W_min": 10, "force_learn_all": True},                      # both register halves of the ring
                {"frame_len": 256},                                                   # S = 2, the 8-lane FFT
                {"smooth_N_E": True, "noise_replenish_from_all_subframes": True, "noise_buffer_ttl_frames": 40,
                 "det.k_subframes": 3, "W": 20, "W_min": 5},                          # replenish + smoothing
                {"frame_len": 1024, "W": 40}]                                         # S = 8 (generic FFT kernel on both sides)
    for extra in variants:
        params = {"sample_rate": 11162, **extra}
        proc = BandNoiseEstimatorProcessor()
        monkeypatch.setenv("APT_BNE_FFT_GENERIC", "1")
        monkeypatch.setenv("APT_BNE_STATE_SERIAL", "1")
        monkeypatch.setenv("APT_BNE_FILTER_SERIAL", "1")
        old = proc.run_batch(clips, params)
        monkeypatch.delenv("APT_BNE_STATE_SERIAL")
        monkeypatch.delenv("APT_BNE_FILTER_SERIAL")
        warp = proc.run_batch(clips, params)          # wavefront filters + warp state machine on the generic FFT: bit-equal
        monkeypatch.delenv("APT_BNE_FFT_GENERIC")
        new = proc.run_batch(clips, params)           # both new kernels
        for (r0, s0), (r1, s1), (r2, s2) in zip(old, warp, new):
            assert r0.keys() == r1.keys() == r2.keys()
            for k, v in r0.items():
                if isinstance(v, float) and np.isnan(v):
                    assert np.isnan(r1[k]) and np.isnan(r2[k]), k
                else:
                    assert r1[k] == v, k
                    assert r2[k] == (pytest.approx(v, rel=1e-9) if isinstance(v, float) else v), k
            for k, v in s0.items():
                if isinstance(v, np.ndarray):
                    assert np.array_equal(s1[k], v, equal_nan=v.dtype.kind == "f"), k
                    if v.dtype.kind == "f":
                        np.testing.assert_allclose(s2[k], v, rtol=1e-9, atol=0, equal_nan=True, err_msg=k)
                    else:
                        assert np.array_equal(s2[k], v), k


def test_legacy_roe_matches_reference(torch_cuda):
    """the legacy RoE detector on the GPU against ten runs of the unmodified reference, in the
    reference's call order (its `max_harmonics` module state carries over): drop / peak counts and the per-frame rain
    status bit-exact, float64 series within 1e-9 relative."""
    import json
    from test_roe_oracle import roe_cases, case_pcm, check_roe
    from audio_processing_tools_b200.edge import dsp_rain_detection as roe
    g, meta, defaults = roe_cases()
    roe.max_harmonics = defaults["num_harmonics"]
    for m in meta:
        sc = json.loads(str(g[m["name"] + "__scalars"]))
        assert roe.max_harmonics == sc["max_harmonics_in"], m["name"]
        params = dict(defaults)
        params.update(m["extra"])
        drops, frain_mean, st = roe.rain_detection_algo(pcm_to_f32(case_pcm(m)), **params)
        check_roe((drops, frain_mean, st, roe.max_harmonics), g, m["name"], rtol=1e-9)


def test_legacy_roe_batch_and_rain_processor(torch_cuda):
    """A ragged int16 batch equals the per-clip calls (module state included), and the function is a drop-in `fn`
    for RainProcessor (processors.py:84-142)."""
    from audio_processing_tools_b200.edge import dsp_rain_detection as roe
    from audio_processing_tools_b200.processors import RainProcessor
    clips = [synth_clip_i16(10.0, 91, 10.0), synth_clip_i16(7.5, 92, 3.0), synth_clip_i16(10.0, 93, 30.0)]
    roe.max_harmonics = 6
    single = [roe.rain_detection_algo(pcm_to_f32(c), **roe.default_params) for c in clips]
    mh = roe.max_harmonics
    roe.max_harmonics = 6
    batch = roe.rain_detection_algo_batch(clips, **roe.default_params)
    assert roe.max_harmonics == mh
    for (d1, f1, s1), (d2, f2, s2) in zip(single, batch):
        assert d1 == d2 and f1 == f2
        for k in ("raining", "kurtosis", "crest_factor", "diff_energy", "Nov0"):
            assert np.array_equal(s1[k], s2[k], equal_nan=True), k
    res, st = RainProcessor(name="rain", fn=roe.rain_detection_algo).run(pcm_to_f32(clips[0]), dict(roe.default_params))
    assert res["rain_drops"] == single[0][0] and res["rain_drop_count"] == single[0][2]["rain_drop_count"]
    assert st["processor"] == "rain"


def test_legacy_roe_filter_kernels_agree(torch_cuda, monkeypatch):
    """The RoE filter cascade with several parts per warp (one shuffle per step) against the one-warp-per-part kernel
    it replaces (APT_ROE_FILTER_SERIAL=1): same operations on the same values, every output bit-equal."""
    from audio_processing_tools_b200.edge import dsp_rain_detection as roe
    clips = [synth_clip_i16(10.0, 191, 10.0), synth_clip_i16(7.5, 192, 3.0), synth_clip_i16(3.1, 193, 30.0),
             synth_clip_i16(10.0, 194, 0.0), synth_clip_i16(5.9, 195, 3.0)]
    roe.max_harmonics = 6
    monkeypatch.setenv("APT_ROE_FILTER_SERIAL", "1")
    old = roe.rain_detection_algo_batch(clips, **roe.default_params)
    mh = roe.max_harmonics
    monkeypatch.delenv("APT_ROE_FILTER_SERIAL")
    roe.max_harmonics = 6
    new = roe.rain_detection_algo_batch(clips, **roe.default_params)
    assert roe.max_harmonics == mh
    for (d1, f1, s1), (d2, f2, s2) in zip(old, new):
        assert d1 == d2 and f1 == f2 and s1.keys() == s2.keys()
        for k, v in s1.items():
            if isinstance(v, np.ndarray):
                assert np.array_equal(v, s2[k], equal_nan=v.dtype.kind == "f"), k
            else:
                assert v == s2[k], k


def test_rain_processor_run_batch_uses_the_gpu_batch(torch_cuda):
    """RainProcessor.run_batch over the legacy RoE fn: one GPU pass for the list, same results as run() per file."""
    from audio_processing_tools_b200.edge import dsp_rain_detection as roe
    from audio_processing_tools_b200.processors import RainProcessor
    clips = [pcm_to_f32(synth_clip_i16(10.0, 91 + i, 10.0)) for i in range(3)]
    proc = RainProcessor(name="rain", fn=roe.rain_detection_algo)
    roe.max_harmonics = 6
    single = [proc.run(c, dict(roe.default_params)) for c in clips]
    roe.max_harmonics = 6
    batch = proc.run_batch(clips, dict(roe.default_params))
    for (r1, s1), (r2, s2) in zip(single, batch):
        assert r1["rain_drops"] == r2["rain_drops"] and r1["rain_drop_count"] == r2["rain_drop_count"]
        assert np.array_equal(s1["raining"], s2["raining"])


def test_adaptive_q_across_time_segments(torch_cuda, oracle_mod):
    """adaptive_q_enable on clips long enough for several time segments (the clip's rain EMA is carried between the
    segments of the pipelined run): the suppressor's noise PSD equals the oracle's bit for bit, ragged batch."""
    from audio_processing_tools_b200.synth import default_params, synth_clip_i16
    params = default_params(check_duration=60, adaptive_q_enable=True, adaptive_q_min=0.08, adaptive_q_alpha=0.97)
    clips = [synth_clip_i16(95.0, 71, 10.0), synth_clip_i16(61.5, 72, 3.0), synth_clip_i16(140.0, 73, 0.5)]
    eng = make_engine(params)
    plan, out = eng.run_clips(clips, ("noise_psd",))
    for c, pcm in enumerate(clips):
        f0, f1 = int(plan.frame_off[c]), int(plan.frame_off[c + 1])
        m, s = oracle_mod.run(pcm_to_f32(pcm), dict(params, keep_state_debug=True))
        assert np.array_equal(out["frame_class"][f0:f1], s["frame_class"])
        assert np.array_equal(out["noise_psd"][f0:f1], s["N2_band"])
        assert out["clip_stats"][c][6] == pytest.approx(m["mean_noise_floor_db"], rel=1e-6)
        assert out["clip_stats"][c][7] == pytest.approx(m["median_noise_floor_db"], rel=1e-6)
    # and it is not the fixed-q result
    eng0 = make_engine(default_params(check_duration=60))
    _, out0 = eng0.run_clips(clips, ("noise_psd",))
    assert not np.array_equal(out0["noise_psd"], out["noise_psd"])
    eng.close()
    eng0.close()


@pytest.mark.parametrize("extra", [{"pre_smooth_frames": 3}, {"median_frames": 5}, {"median_frames": 8, "pre_smooth_frames": 16},
                                   {"pre_smooth_frames": 4, "median_frames": 4, "adaptive_q_enable": True}])
def test_tracker_smoothing_options_across_time_segments(torch_cuda, oracle_mod, extra):
    """pre_smooth_frames / median_frames on clips long enough for several time segments (cumulative sums, ring and median
    windows cross the segment borders), ragged batch: labels, both passes' noise PSD planes and the lagged baseline equal
    the oracle's bit for bit (the oracle equals the reference on the alt_*_presmooth / median fixtures)."""
    from audio_processing_tools_b200.synth import default_params, synth_clip_i16
    params = default_params(check_duration=60, **extra)
    clips = [synth_clip_i16(95.0, 81, 10.0), synth_clip_i16(61.5, 82, 3.0), synth_clip_i16(130.0, 83, 0.5)]
    eng = make_engine(params)
    plan, out = eng.run_clips(clips, ("noise_psd", "det_noise_psd", "det_noise_lag"))
    for c, pcm in enumerate(clips):
        f0, f1 = int(plan.frame_off[c]), int(plan.frame_off[c + 1])
        m, s = oracle_mod.run(pcm_to_f32(pcm), dict(params, keep_state_debug=True))
        assert np.array_equal(out["det_noise_psd"][f0:f1], s["N1_band"])
        assert np.array_equal(out["det_noise_lag"][f0:f1], s["Nlag_band"])
        assert np.array_equal(out["frame_class"][f0:f1], s["frame_class"])
        assert np.array_equal(out["noise_psd"][f0:f1], s["N2_band"])
        assert out["clip_stats"][c][7] == pytest.approx(m["median_noise_floor_db"], rel=1e-6)
    # default flags (mode-bin lanes only, no debug planes): same labels
    plan2, out2 = eng.run_clips(clips, ())
    assert np.array_equal(out2["frame_class"], out["frame_class"])
    assert np.array_equal(out2["clip_stats"], out["clip_stats"])
    eng.close()
