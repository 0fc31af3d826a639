"""Per-kernel share of one bench step from an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of bench.py.
usage: python profiles/launch_shares.py launches.csv > summary.txt
Only the full-batch launches (largest grid per kernel) are used: the e2e leg of bench.py adds 24 clip-group launches per
step whose serial kernels (trk1, base, trk2) cost the same chain latency at 42 clips as at 1000."""
import collections, csv, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
L = []
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    n = re.sub(r"<.*", "", r[ix["Kernel Name"]].split("(")[0]).replace("void ", "").replace("apt::", "")
    if n.startswith("at::"):
        continue
    v = float(r[ix["Metric Value"]]) * {"ns": 1e-6, "us": 1e-3, "ms": 1}.get(r[ix["Metric Unit"]], 1e-6)
    gs = 1
    for x in re.findall(r"\d+", r[ix["Grid Size"]]):
        gs *= int(x)
    L.append((n, gs, v))
mx = collections.defaultdict(int)
for n, gs, v in L:
    mx[n] = max(mx[n], gs)
full = collections.defaultdict(list)
for n, gs, v in L:
    if gs == mx[n]:
        full[n].append(v)
mult = {"select_hist_kernel": 2, "select_scan_kernel": 3}
tot = sum(sum(v) / len(v) * mult.get(n, 1) for n, v in full.items())
print("ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 1  (1 B200)")
print("Full-batch (1000-clip) launches only; cold-cache, serialised per-launch times: compare the SHARE per kernel with")
print("bench.py's kernel_ms_per_step.\n")
for n, v in sorted(full.items(), key=lambda kv: -sum(kv[1]) / len(kv[1]) * mult.get(kv[0], 1)):
    per = sum(v) / len(v)
    print("%-22s full-batch launches %3d  mean %8.3f ms  x%d per step  share %5.1f%%" % (n, len(v), per, mult.get(n, 1), 100 * per * mult.get(n, 1) / tot))
print("sum per step %.1f ms" % tot)
