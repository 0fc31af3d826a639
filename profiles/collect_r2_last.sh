#!/bin/bash
# Last evidence of round 2 with the final code: bash profiles/collect_r2_last.sh   (outputs under gpurun_out/r2last/)
set -u
O=gpurun_out/r2last; mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/pytest_gpu.txt 2>&1; tail -2 $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
python bench.py > $O/bench_default.out 2> $O/bench_default.err; tail -1 $O/bench_default.out > $O/bench_default.json
python bench.py --impl reference > $O/bench_reference.out 2> $O/bench_reference.err; tail -1 $O/bench_reference.out > $O/bench_reference.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2last/bench_default.json").read())
print("ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
print(d["roofline"]["kernel_ms_per_step"]); print(d["roofline"]["frac"], d["roofline"]["achieved"])
r = json.loads(open("gpurun_out/r2last/bench_reference.json").read()); print("reference", r["value"])
PY
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > /dev/null 2>&1
python profiles/launch_shares.py $O/launches_bench.csv > $O/launches_bench_summary.txt 2>&1; head -14 $O/launches_bench_summary.txt
