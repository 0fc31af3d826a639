"""BASELINE configs 2 and 5: features stage (STFT + power + band energies) on one 1-hour synthetic clip,
per frame size / hop, float64 and float32 FFT.  Device-timed (CUDA events, 3 warm-ups, 5 timed passes);
inputs resident in HBM; an L2 flush (256 MB write) precedes every timed pass because one clip (80 MB) fits L2.
Prints one JSON line per configuration."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_processing_tools_b200.config import build_noise_config
from audio_processing_tools_b200.engine import BatchEngine
from audio_processing_tools_b200.synth import default_params, synth_clip_i16

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 3600.0
only = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else None      # one geometry (for an ncu capture)
only_fft = sys.argv[4] if len(sys.argv) > 4 else None
pcm_h = synth_clip_i16(seconds, 77, 3.0)
pcm = torch.from_numpy(pcm_h).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
for n_fft, hop in ((256, 128), (256, 64), (512, 256), (512, 128), (1024, 512), (1024, 256), (2048, 1024), (2048, 512), (4096, 2048), (4096, 1024)):
    if only and (n_fft, hop) != only:
        continue
    for fft in ("f64", "f32", "tc"):
        if only_fft and fft != only_fft:
            continue
        if fft == "tc" and (n_fft, hop) != (256, 128):
            continue            # the DFT-as-GEMM variant exists for the short frame only (its work grows with n_fft^2)
        for planes in (("band_energy",), ("band_energy", "S")):
            if fft == "tc" and "S" in planes:
                continue
            params = default_params(check_duration=seconds, n_fft=n_fft, hop=hop)
            eng = BatchEngine(build_noise_config(11162, params), 11162, fft_f64={"f64": True, "f32": False, "tc": "tc"}[fft])
            plan = eng.plan_for([pcm_h.size])
            bufs = eng.alloc_outputs(plan, planes, full=False)
            for _ in range(3):
                eng.run_device(plan, pcm, bufs, full=False)
            torch.cuda.synchronize()
            ms = []
            for _ in range(5):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); eng.run_device(plan, pcm, bufs, full=False); e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            t = float(np.median(ms)) * 1e-3
            M = eng.rp.M
            bytes_algo = pcm_h.size * 2 + plan.nF * ((M + 1) * 4 + (8 * eng.rp.F if "S" in planes else 0))
            print(json.dumps({"workload": "features stage, 1 clip x %gs" % seconds, "n_fft": n_fft, "hop": hop, "fft": fft,
                              "outputs": "+".join(planes), "frames": plan.nF, "ms": t * 1e3, "audio_s_per_s": seconds / t,
                              "bytes_algo": bytes_algo, "achieved_gbs": bytes_algo / t / 1e9, "frac_of_measured_hbm": bytes_algo / t / 1e9 / peak}), flush=True)
            eng.close()
