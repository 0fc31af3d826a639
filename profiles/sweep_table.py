"""Markdown table of the frame-size sweep (BASELINE configs[4]): device-timed features stage on one 1-hour clip
(profiles/sweep_bench.py -> sweep_features_1h_r2.jsonl) beside the ncu pass of each geometry (profiles/r2/sweep/*.csv).
usage: python profiles/sweep_table.py > profiles/r2/sweep/TABLE.md"""
import csv, glob, json, os
here = os.path.dirname(os.path.abspath(__file__))
rows = [json.loads(l) for l in open(os.path.join(here, "r2", "sweep_features_1h_r2.jsonl"))]
def ncu(path):
    out = {}
    for r in csv.reader(open(path)):
        if len(r) > 14 and r[0] != "ID":
            out[r[12]] = (r[14], r[13]); out["kernel"] = (r[4].split("(")[0].replace("void ", ""), "")
    return out
print("| n_fft / hop | band bins K | float64 FFT ms | float32 FFT ms | tensor-core DFT ms | best audio-s/s | ncu kernel | ncu us | DRAM MB | issue % | fp64 pipe % | tensor pipe % |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
geoms = sorted({(r["n_fft"], r["hop"]) for r in rows}, key=lambda g: (g[0], -g[1]))
for n_fft, hop in geoms:
    sel = {r["fft"]: r for r in rows if (r["n_fft"], r["hop"]) == (n_fft, hop) and r["outputs"] == "band_energy"}
    K = int(round(71 * n_fft / 256))
    best = max(sel.values(), key=lambda r: r["audio_s_per_s"])
    for tag, fn in (("f64", f"stft_{n_fft}_{hop}_f64.csv"), ("tc", f"tcdft_{n_fft}_{hop}.csv")):
        p = os.path.join(here, "r2", "sweep", fn)
        if not os.path.exists(p):
            continue
        m = ncu(p)
        g = lambda k: m.get(k, ("n/a", ""))[0]
        mb = (float(g("dram__bytes_read.sum")) + float(g("dram__bytes_write.sum"))) / 1e6
        print(f"| {n_fft} / {hop} | {K} | {sel['f64']['ms']:.3f} | {sel['f32']['ms']:.3f} | " + (f"{sel['tc']['ms']:.3f}" if 'tc' in sel else "-") +
              f" | {best['audio_s_per_s'] / 1e6:.2f} M ({best['fft']}) | {m['kernel'][0]} | {float(g('gpu__time_duration.sum')) / 1e3:.1f} | {mb:.1f} | "
              f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active')} | {g('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active')} | {g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')} |")
