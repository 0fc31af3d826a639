set -u
T=/tmp/ncu_stft; mkdir -p $T gpurun_out/r2stft
APT_SEGMENTS=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:"stft256_kernel" -c 1 -o $T/stft python profiles/run_small.py 16 600 > /dev/null 2>&1
ncu -i $T/stft.ncu-rep --page source --csv > $T/src.csv 2>/dev/null
python profiles/sass_mix.py $T/src.csv > gpurun_out/r2stft/sass_mix_stft256.txt 2>&1
python - <<'PY'
import csv, collections, re
rows = list(csv.reader(open("/tmp/ncu_stft/src.csv")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
out = open("gpurun_out/r2stft/stft256_sass_executed.txt", "w")
keep = [k for k in ("Address", "Source", "Instructions Executed", "# Samples", "Thread Instructions Executed") if k in ix]
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    out.write("\t".join(r[ix[k]] for k in keep) + "\n")
out.close()
PY
wc -l gpurun_out/r2stft/*; head -30 gpurun_out/r2stft/sass_mix_stft256.txt
