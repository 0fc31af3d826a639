"""Per-kernel share of one step from an ncu launch list (ncu --metrics gpu__time_duration.sum --csv) of bench.py.
The pipelined run launches most kernels once per time segment, so durations are SUMMED per kernel and divided by the
number of steps in the capture (= launches of finalize_kernel).  Per-launch times under ncu are cold-cache and serialised:
compare the SHARE per kernel with bench.py's kernel_ms_per_step, not the absolute.
usage: python profiles/launch_shares.py launches.csv"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0] != "ID"]
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "").replace("apt::", "").split("<")[0].strip()
    if name.startswith("at::") or not name:
        continue
    tot[name] += float(r[14]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[13], 1e-6)
    cnt[name] += 1
steps = max(1, cnt.get("finalize_kernel", 1))
total = sum(tot.values()) / steps
print(f"{sys.argv[1]}: {len(rows)} launches, {steps} steps; per-step sums (ms), launches per step, share")
for name, ms in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{name:24s} {ms / steps:8.3f} ms  {cnt[name] / steps:6.1f} launches  {100 * ms / steps / total:5.1f} %")
print(f"{'sum per step':24s} {total:8.3f} ms")
