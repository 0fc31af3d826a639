#!/bin/bash
# Band noise estimator after the warp-per-clip state machine and the 16-lane FFT: bash profiles/collect_r2_bne.sh
# (outputs under gpurun_out/r2bne/)
set -u
O=gpurun_out/r2bne; mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
B="python bench.py --workload bne --steps 3 --warmup 3"
$B 2> $O/bne_default.err | tail -1 > $O/bench_bne.json
APT_BNE_STATE_SERIAL=1 APT_BNE_FFT_GENERIC=1 APT_BNE_FILTER_SERIAL=1 $B --no-cpu 2> /dev/null | tail -1 > $O/bench_bne_old_kernels.json
APT_BNE_FFT_GENERIC=1 $B --no-cpu 2> /dev/null | tail -1 > $O/bench_bne_generic_fft.json
APT_BNE_FILTER_SERIAL=1 $B --no-cpu 2> /dev/null | tail -1 > $O/bench_bne_serial_filter.json
for s in 4 8 32 64; do APT_BNE_SEG=$s $B --no-cpu 2> /dev/null | tail -1 > $O/bench_bne_seg$s.json; done
for f in $O/bench_bne*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]) if d.get("e2e") else None)
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_bne.csv $B --no-cpu > /dev/null 2>&1
grep -c bne $O/launches_bne.csv
