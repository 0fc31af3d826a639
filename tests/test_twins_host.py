"""Host-side logic of the neighbouring engines' Python twins (no GPU): configuration resolution and part tables."""
import math

import numpy as np
import pytest

from audio_processing_tools_b200.edge import dsp_rain_detection as roe
from audio_processing_tools_b200.edge.band_noise_estimator import BandNoiseEstimatorConfig, hz_to_bin
from audio_processing_tools_b200.edge.band_noise_processor import BandNoiseEstimatorProcessor


def test_roe_configure_parameters_matches_reference_defaults():
    cfg = roe.configure_parameters(**roe.default_params)
    # dsp_rain_detection.py:1326-1365: frame 2^ceil(log2(11162/45)) = 256, hop 2^ceil(log2(111.62)) = 128, M = 20
    assert (cfg["frame_length"], cfg["hop_length"], cfg["M"]) == (256, 128, 20)
    assert cfg["rain_thr_hn"] == pytest.approx(4.5 + 4.0 + 3.5)
    with pytest.raises(TypeError):
        roe.configure_parameters(not_a_parameter=1)               # as the reference's keyword signature does
    with pytest.raises(NotImplementedError):
        roe.configure_parameters(freq_resolution=20)              # frame_length 1024
    with pytest.raises(NameError):
        roe.configure_parameters(nf=1)                            # the reference calls an undefined function there


def test_roe_part_table_follows_the_two_second_walk():
    cfg = roe.configure_parameters(**roe.default_params)
    fs = roe.FS_ANALYSIS
    lens = [10 * fs, int(7.3 * fs), fs // 2]
    clip, start, plen, last_ok = roe._part_table(lens, cfg)
    # clip 0: five full parts of int(256 * (2 * 11162 / 256)) samples at offsets int(11162 * k)
    assert list(clip[:5]) == [0] * 5 and list(plen[:5]) == [2 * fs] * 5
    assert list(start[:5]) == [fs * k for k in (0, 2, 4, 6, 8)]
    # clip 1 (7.3 s): parts at 0, 2, 4 s are full, the one at 6 s has 1.3 s, the one at 8 s is empty and dropped
    c1 = np.flatnonzero(clip == 1)
    assert len(c1) == 4 and list(start[c1] - lens[0]) == [0, 2 * fs, 4 * fs, 6 * fs]
    assert plen[c1[-1]] == lens[1] - 6 * fs
    # clip 2 is shorter than one second: no part at all
    assert not np.any(clip == 2)
    assert last_ok == [True, False, False]
    assert int(plen.max()) // 128 + 2 <= 256        # frame slots of a part fit one CTA of the part kernels


def test_band_noise_config_resolution():
    proc = BandNoiseEstimatorProcessor()
    cfg = proc._build_config({"sample_rate": 11162, "det.M_db": 4.0, "W": 20, "W_min": 5, "dtype": "float64"})
    assert isinstance(cfg, BandNoiseEstimatorConfig) and cfg.det.M_db == 4.0 and cfg.det.n_fft == cfg.frame_len == 512
    P = proc._resolve(cfg)
    assert (P.N, P.sub_len, P.S) == (512, 128, 4) and P.ns_h == 2 and P.ns_b == 4
    assert P.warm > 0 and P.warm % 512 == 0
    assert (P.prim_b0, P.prim_b1) == (hz_to_bin(450.0, 11162, 512), hz_to_bin(650.0, 11162, 512))
    freqs = np.fft.rfftfreq(512, 1 / 11162)
    idx = np.flatnonzero((freqs >= 400.0) & (freqs <= 700.0))
    assert (P.mask_b0, P.mask_b1) == (idx[0], idx[-1])
    assert P.M_ratio == pytest.approx(10 ** 0.4)
    with pytest.raises(ValueError):
        proc._build_config({"q": 1.5})
    with pytest.raises(NotImplementedError):
        proc._resolve(proc._build_config({"dtype": "float32"}))
    with pytest.raises(NotImplementedError):
        proc._resolve(proc._build_config({"subhop": 64}))


def test_band_noise_rejects_hop_other_than_frame():
    with pytest.raises(ValueError, match="hop == frame_len"):
        BandNoiseEstimatorProcessor().run_batch([np.zeros(4096, np.float32)], {"hop": 256})


def test_rain_processor_batch_hook_uses_the_companion():
    """RainProcessor.run_batch: one call of fn.batch for the list, the same packaging as run() per file."""
    from audio_processing_tools_b200.processors import RainProcessor
    calls = []

    def fn(audio, **params):
        return int(audio.size), 1.5, {"rain_drop_count": int(audio.size), "extra": params.get("k")}

    def fn_batch(audios, **params):
        calls.append(len(audios))
        return [fn(a, **params) for a in audios]

    clips = [np.zeros(n, np.float32) for n in (20, 30)]
    params = {"sample_rate": 10, "check_duration": 2, "k": 7}
    plain = RainProcessor(name="rain", fn=fn).run_batch(clips, params)          # no companion: per-file run()
    fn.batch = fn_batch
    batched = RainProcessor(name="rain", fn=fn).run_batch(clips, params)
    assert calls == [2]
    for (r1, s1), (r2, s2) in zip(plain, batched):
        assert {k: v for k, v in r1.items() if k != "latency_s"} == {k: v for k, v in r2.items() if k != "latency_s"}
        assert s1["processor"] == s2["processor"] == "rain" and s1["extra"] == s2["extra"] == 7
    with pytest.raises(ValueError):
        RainProcessor(name="rain", fn=fn).run_batch([np.zeros(5, np.float32)], params)     # shorter than check_duration
    assert roe.rain_detection_algo.batch is roe.rain_detection_algo_batch


def test_band_noise_packaging_from_batch_arrays():
    """Host side of BandNoiseEstimatorProcessor.run_batch: the per-clip result / state dictionaries are cut out of the
    batch-wide column-major arrays the device returns (no GPU needed: synthetic device outputs).  Medians follow
    np.median (two-element mean for even counts), masks unpack bit s = subframe s, every array is the clip's own copy."""
    from audio_processing_tools_b200 import _lib
    rng = np.random.default_rng(5)
    proc = BandNoiseEstimatorProcessor()
    for n in (1, 2, 5, 6, 1307, 1308):
        a = rng.random(n)
        assert proc._median(a) == float(np.median(a)), n
    params = {"sample_rate": 11162}
    cfg = proc._build_config(params)
    N, S = 512, 4
    nfr = np.array([7, 0, 12, 1], dtype=np.int64)
    nF = int(nfr.sum())
    clips = [np.zeros(int(n) * N + 3, dtype=np.int16) for n in nfr]
    fo_t = rng.random((_lib.BNE_FRAME_F, nF))
    fo_t[11] = (rng.random(nF) > 0.5).astype(np.float64)
    mask = rng.integers(0, 16, nF).astype(np.uint8)
    sub = rng.random((nF, S))
    st = np.arange(len(clips) * _lib.BNE_STATS, dtype=np.float64).reshape(len(clips), _lib.BNE_STATS)
    outs = proc._package_batch(cfg, clips, params, fo_t, mask, sub, st, nfr, S, N, 11162)
    assert len(outs) == len(clips)
    f0 = 0
    for c, (res, state) in enumerate(outs):
        n = int(nfr[c])
        assert res["n_frames"] == n and state["M_band"].shape == (n,) and state["rain_submask"].shape == (n, S)
        if n == 0:
            assert np.isnan(res["gain_med"]) and res["energy_stats__noise_effective_q"] == float(cfg.q)
            continue
        assert np.array_equal(state["M_band"], fo_t[0, f0:f0 + n]) and np.array_equal(state["G_mag"], fo_t[4, f0:f0 + n])
        assert res["gain_med"] == float(np.median(fo_t[4, f0:f0 + n]))
        assert res["fft_rain_frac"] == float(np.mean(fo_t[11, f0:f0 + n] > 0.5))
        assert np.array_equal(state["rain_submask"], ((mask[f0:f0 + n, None] >> np.arange(S)[None, :]) & 1).astype(bool))
        assert np.array_equal(state["N_sub"], np.repeat(fo_t[10, f0:f0 + n, None], S, axis=1))
        assert np.array_equal(state["subE"], sub[f0:f0 + n]) and np.array_equal(state["times_s"], np.arange(n) * N / 11162)
        assert res["energy_stats__noise_frame_count"] == int(st[c, 3]) and res["energy_stats__noise_energy_sum"] == float(st[c, 0])
        state["M_band"][0] = -1.0          # the clip's own copy: the batch array is untouched
        assert fo_t[0, f0] != -1.0
        f0 += n
