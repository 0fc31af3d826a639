#!/bin/bash
# Engines beside the path after the wavefront filters / warp state machine / pinned staging:
# bash profiles/collect_r2_neighbours.sh   (outputs under gpurun_out/r2nb/)
set -u
O=gpurun_out/r2nb; mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
for w in bne roe dsd; do
  B="python bench.py --workload $w --steps 3 --warmup 3"
  $B 2> $O/$w.err | tail -1 > $O/bench_$w.json
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_$w.csv $B --no-cpu > /dev/null 2>&1
done
APT_ROE_FILTER_SERIAL=1 python bench.py --workload roe --steps 3 --warmup 3 --no-cpu 2> /dev/null | tail -1 > $O/bench_roe_serial_filter.json
APT_BNE_STATE_SERIAL=1 APT_BNE_FFT_GENERIC=1 APT_BNE_FILTER_SERIAL=1 python bench.py --workload bne --steps 3 --warmup 3 --no-cpu 2> /dev/null | tail -1 > $O/bench_bne_old_kernels.json
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]) if d.get("e2e") else None,
          "e2e ms", round(d["e2e"]["ms_per_step"], 1) if d.get("e2e") else None)
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
