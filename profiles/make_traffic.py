"""Per-kernel DRAM traffic of one pass of profiles/run_small.py from an `ncu --set full` report.
usage: python profiles/make_traffic.py report.ncu-rep frames out.json
The report may hold several passes; the LAST pass (last launch group of every kernel) is used."""
import collections, csv, io, json, subprocess, sys
rep, frames, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def to_bytes(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
def to_us(v, u):
    return float(v) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}[u]
def base(name):
    n = name.split("(")[0].split("<")[0].split()[-1]
    return n.replace("apt::", "")
launches = [(base(r[ix["Kernel Name"]]), r) for r in rows[2:]]
launches = [(n, r) for n, r in launches if not n.startswith("at::")]   # torch's own fill / reduce kernels of the driver script
n_pass = max(1, sum(1 for n, _ in launches if n == "stft256_kernel"))
per_kernel = collections.Counter(n for n, _ in launches)
seen = collections.Counter()
k = collections.OrderedDict()
for n, r in launches:
    seen[n] += 1
    if seen[n] <= per_kernel[n] - per_kernel[n] // n_pass:   # keep the last pass only
        continue
    d = k.setdefault(n, {"launches_per_pass": 0, "dram_bytes_per_pass": 0.0, "ncu_us_per_pass": 0.0})
    d["launches_per_pass"] += 1
    d["dram_bytes_per_pass"] += to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) + to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
    d["ncu_us_per_pass"] += to_us(r[ix["gpu__time_duration.sum"]], units[ix["gpu__time_duration.sum"]])
for d in k.values():
    d["dram_bytes_per_frame"] = d["dram_bytes_per_pass"] / frames
json.dump({"workload": (sys.argv[4] if len(sys.argv) > 4 else "profiles/run_small.py (full pipeline, one pass, one time segment)"), "frames": frames, "passes_in_report": n_pass,
           "how": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per kernel; per_frame = bytes / frames",
           "kernels": k}, open(out, "w"), indent=1)
print(json.dumps({n: round(d["dram_bytes_per_frame"], 1) for n, d in k.items()}))
