"""Aggregate an `ncu --page source --csv` dump: executed warp-instructions and stall samples per SASS opcode.
usage: python profiles/sass_mix.py dump.csv [frames_x_warps]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); samp = collections.Counter(); tot = 0; stot = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    src = r[ix["Source"]].strip()
    try:
        n = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
    except ValueError:
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src[:10]
    base = op.split(".")[0]
    if base in ("F2F", "I2F", "F2I", "MUFU"):
        base = ".".join(op.split(".")[:3])
    ops[base] += n; samp[base] += s; tot += n; stot += s
div = float(sys.argv[2]) if len(sys.argv) > 2 else None
print("total warp-instructions", tot, "samples", stot, ("per unit %.1f" % (tot / div)) if div else "")
for op, n in ops.most_common(45):
    print("%-16s %10d %5.1f%%  samples %6d %5.1f%%" % (op, n, 100 * n / tot, samp[op], 100 * samp[op] / max(1, stot)) + (("  per unit %.2f" % (n / div)) if div else ""))
