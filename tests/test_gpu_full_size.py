"""Parity at BASELINE.json's full sizes through size-independent properties (GPU).

config 3: 1,000 x 10-min clips, full pipeline -- the batch tiles 16 distinct synthetic clips, so
  * every copy of a clip must give identical labels / events / statistics wherever it sits in the batch
    (batch invariance: no cross-clip state),
  * the distinct clips are checked against the CPU oracle at full clip length (bit-exact labels and events),
  * event lists are exactly the ascending positions of the RAIN labels and counts add up (compaction),
  * a second run is bit-identical (determinism / idempotence of the plan's scratch reuse).
config 2: one 1-hour clip, features stage -- exact binary scaling: S(x/2) = S(x)/2 and P(x/2) = P(x)/4
  bit for bit (a power-of-two factor commutes with every rounding in the float64 FFT and the float32 power).
"""
import hashlib

import numpy as np
import pytest

from audio_processing_tools_b200.synth import batch_clip_spec, default_params, pcm_to_f32, synth_clip_i16

pytestmark = pytest.mark.gpu

N_CLIPS, SECONDS, N_BASE = 1000, 600.0, 16


def _engine(params, **kw):
    from audio_processing_tools_b200.config import build_noise_config
    from audio_processing_tools_b200.engine import BatchEngine
    return BatchEngine(build_noise_config(11162, params), 11162, **kw)


def test_config3_full_batch_properties(oracle_mod):
    import torch
    params = default_params(check_duration=SECONDS)
    eng = _engine(params)
    base = [synth_clip_i16(SECONDS, *batch_clip_spec(5000 + i)) for i in range(N_BASE)]
    N = base[0].size
    plan = eng.plan_for([N] * N_CLIPS)
    T = 1 + N // 128
    assert plan.nF == N_CLIPS * T
    dev = torch.device("cuda", 0)
    order = np.random.default_rng(3).integers(0, N_BASE, N_CLIPS)       # which distinct clip sits at each position
    order[:N_BASE] = np.arange(N_BASE)
    base_dev = torch.from_numpy(np.stack(base)).to(dev)
    pcm = base_dev[torch.from_numpy(order).to(dev)].reshape(-1).contiguous()
    del base_dev
    bufs = eng.alloc_outputs(plan, (), full=True)
    eng.run_device(plan, pcm, bufs, full=True)
    torch.cuda.synchronize()
    out = {k: v.cpu().numpy() for k, v in bufs.items()}
    # determinism: a second pass over the same plan (scratch reused) is bit-identical
    eng.run_device(plan, pcm, bufs, full=True)
    torch.cuda.synchronize()
    for k, v in bufs.items():
        if k == "event_idx":
            continue        # only the first event_count entries per clip are defined
        assert np.array_equal(v.cpu().numpy(), out[k]), k
    fc = out["frame_class"].reshape(N_CLIPS, T)
    ev = out["event_idx"].reshape(N_CLIPS, T)
    cnt = out["event_count"]
    stats = out["clip_stats"]
    # compaction: events are the ascending RAIN positions, counts add up
    assert np.array_equal(cnt, (fc == 2).sum(axis=1))
    for c in range(0, N_CLIPS, 37):
        assert np.array_equal(ev[c, :cnt[c]], np.flatnonzero(fc[c] == 2))
    assert np.array_equal(stats[:, 0], np.arange(N_CLIPS, dtype=np.float32))
    assert np.array_equal(stats[:, 1], cnt.astype(np.float32))
    assert np.array_equal(out["rain_conf"].reshape(N_CLIPS, T), (fc == 2).astype(np.float32))
    assert np.array_equal(out["noise_conf"].reshape(N_CLIPS, T), (fc != 2).astype(np.float32))
    # batch invariance: all copies of a distinct clip agree bit for bit
    digest = {}
    for c in range(N_CLIPS):
        h = hashlib.sha1(fc[c].tobytes() + stats[c, 1:].tobytes()).hexdigest()
        assert digest.setdefault(int(order[c]), h) == h, f"clip at position {c} differs from its other copies"
    assert len(set(digest.values())) == N_BASE
    # full-length parity with the oracle on the distinct clips
    fcs, counts = oracle_mod.process_batch_i16(base, params, n_threads=8)
    for i in range(N_BASE):
        assert np.array_equal(fc[i], fcs[i]), f"labels of distinct clip {i} differ from the oracle"
        assert int(cnt[i]) == int(counts[i])
    m, s = oracle_mod.run(pcm_to_f32(base[3]), params)
    assert stats[3, 6] == pytest.approx(m["mean_noise_floor_db"], rel=1e-6)
    assert stats[3, 7] == m["median_noise_floor_db"]
    eng.close()


def test_config2_one_hour_clip_scaling():
    import torch
    seconds = 3600.0
    params = default_params(check_duration=seconds)
    eng = _engine(params)
    x = pcm_to_f32(synth_clip_i16(seconds, 77, 3.0))
    plan = eng.plan_for([x.size])
    assert plan.nF == 313932
    dev = torch.device("cuda", 0)
    res = []
    for scale in (1.0, 0.5):
        bufs = eng.alloc_outputs(plan, ("S", "P", "band_energy"), full=False)
        eng.run_device(plan, torch.from_numpy((x * np.float32(scale))).to(dev), bufs, full=False)
        torch.cuda.synchronize()
        res.append({k: v.cpu().numpy() for k, v in bufs.items()})
    assert np.array_equal(res[1]["S"], res[0]["S"] * np.float32(0.5))
    assert np.array_equal(res[1]["P"], res[0]["P"] * np.float32(0.25))
    # band energies add eps before rounding: equal up to that
    np.testing.assert_allclose(res[1]["band_energy"][:5], res[0]["band_energy"][:5] * np.float32(0.25), rtol=1e-6, atol=0)
    # Parseval-style sanity on a sample of frames: sum |S|^2 over the (Hermitian-completed) spectrum = 256 * sum (w x)^2
    import scipy.signal
    w = scipy.signal.get_window("hann", 256, fftbins=True)
    xp = np.pad(x.astype(np.float64), (128, 128))
    P = res[0]["P"].astype(np.float64)
    for t in (0, 1, 777, 150000, 313931):
        fr = w * xp[t * 128:t * 128 + 256]
        lhs = P[t, 0] + P[t, 128] + 2.0 * P[t, 1:128].sum()
        assert lhs == pytest.approx(256.0 * float((fr * fr).sum()), rel=1e-5)
    eng.close()


@pytest.mark.parametrize("n_fft,hop,extra", [(1024, 256, {}), (256, 64, {}),
                                             (256, 128, {"pre_smooth_frames": 3, "median_frames": 5, "adaptive_q_enable": True})])
def test_config5_full_pipeline_other_geometries_at_size(oracle_mod, n_fft, hop, extra):
    """BASELINE configs[4] with the WHOLE pipeline at size: 96 x 10-min clips (6 distinct) at another frame size / with the
    optional tracker smoothing.  Batch invariance (every copy of a clip agrees bit for bit wherever it sits), compaction,
    determinism of a second pass, and full-length parity of the distinct clips with the oracle."""
    import torch
    n_clips, n_base = 96, 6
    params = default_params(check_duration=SECONDS, n_fft=n_fft, hop=hop, **extra)
    eng = _engine(params)
    base = [synth_clip_i16(SECONDS, *batch_clip_spec(7000 + i)) for i in range(n_base)]
    N = base[0].size
    plan = eng.plan_for([N] * n_clips)
    T = 1 + N // hop
    assert plan.nF == n_clips * T
    dev = torch.device("cuda", 0)
    order = np.random.default_rng(5).integers(0, n_base, n_clips)
    order[:n_base] = np.arange(n_base)
    base_dev = torch.from_numpy(np.stack(base)).to(dev)
    pcm = base_dev[torch.from_numpy(order).to(dev)].reshape(-1).contiguous()
    del base_dev
    bufs = eng.alloc_outputs(plan, (), full=True)
    eng.run_device(plan, pcm, bufs, full=True)
    torch.cuda.synchronize()
    out = {k: v.cpu().numpy() for k, v in bufs.items()}
    eng.run_device(plan, pcm, bufs, full=True)
    torch.cuda.synchronize()
    for k in ("frame_class", "rain_conf", "noise_conf", "event_count", "clip_stats"):
        assert np.array_equal(bufs[k].cpu().numpy(), out[k]), k
    fc = out["frame_class"].reshape(n_clips, T)
    ev = out["event_idx"].reshape(n_clips, T)
    cnt = out["event_count"]
    assert np.array_equal(cnt, (fc == 2).sum(axis=1))
    for c in range(0, n_clips, 11):
        assert np.array_equal(ev[c, :cnt[c]], np.flatnonzero(fc[c] == 2))
    digest = {}
    for c in range(n_clips):
        h = hashlib.sha1(fc[c].tobytes() + out["clip_stats"][c, 1:].tobytes()).hexdigest()
        assert digest.setdefault(int(order[c]), h) == h, f"clip at position {c} differs from its other copies"
    fcs, counts = oracle_mod.process_batch_i16(base, params, n_threads=6)
    for i in range(n_base):
        assert np.array_equal(fc[i], fcs[i]), f"labels of distinct clip {i} differ from the oracle"
        assert int(cnt[i]) == int(counts[i])
    eng.close()
