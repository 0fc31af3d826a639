#!/bin/bash
# Closing evidence of round 2 (second session), run on the GPU box: bash profiles/collect_r2_closing.sh
# Outputs under gpurun_out/r2c/ (text only: the ncu reports stay on the box, their summaries come back).
set -u
O=gpurun_out/r2c; mkdir -p $O; T=/tmp/ncu_r2c; mkdir -p $T
python -m pytest tests -x -q -m gpu > $O/pytest_gpu.txt 2>&1; tail -2 $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
python bench.py > $O/bench_default.out 2> $O/bench_default.err; tail -1 $O/bench_default.out > $O/bench_default.json
python bench.py --impl reference > $O/bench_reference.out 2> $O/bench_reference.err; tail -1 $O/bench_reference.out > $O/bench_reference.json
for w in bne roe dsd; do
  B="python bench.py --workload $w --steps 3 --warmup 3"
  $B 2> $O/$w.err | tail -1 > $O/bench_$w.json
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_$w.csv $B --no-cpu > /dev/null 2>&1
done
APT_ROE_FILTER_SERIAL=1 python bench.py --workload roe --steps 3 --warmup 3 --no-cpu 2> /dev/null | tail -1 > $O/bench_roe_old_filter.json
APT_BNE_STATE_SERIAL=1 APT_BNE_FFT_GENERIC=1 APT_BNE_FILTER_SERIAL=1 python bench.py --workload bne --steps 3 --warmup 3 --no-cpu 2> /dev/null | tail -1 > $O/bench_bne_old_kernels.json
APT_DSD_STATE_SERIAL=1 APT_DSD_FFT_GENERIC=1 python bench.py --workload dsd --steps 3 --warmup 3 --no-cpu 2> /dev/null | tail -1 > $O/bench_dsd_old_kernels.json
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print(sys.argv[1], "ms", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(e["value"]) if e else None, "e2e ms", round(e.get("ms_per_step", 0), 1) if e else None)
except Exception as ex:
    print(sys.argv[1], "unreadable", ex)
PY
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"roe_filter_wave|roe_frame|roe_part" -c 3 -o $T/roe python bench.py --workload roe --steps 1 --warmup 0 --no-cpu > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"bne_filter_wave|bne_state_warp|bne_fft512" -c 3 -o $T/bne python bench.py --workload bne --steps 1 --warmup 0 --no-cpu > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"dsd_fft512|dsd_minutes_warp|dsd_times" -c 3 -o $T/dsd python bench.py --workload dsd --steps 1 --warmup 0 --no-cpu > /dev/null 2>&1
for w in roe bne dsd; do python profiles/ncu_summary.py $T/$w.ncu-rep > $O/ncu_summary_$w.txt 2>&1; done
ncu -i $T/bne.ncu-rep --page source --csv -k regex:bne_filter_wave > $T/s1.csv 2>/dev/null; python profiles/sass_mix.py $T/s1.csv > $O/sass_mix_bne_filter_wave.txt 2>&1
ncu -i $T/bne.ncu-rep --page source --csv -k regex:bne_state_warp > $T/s2.csv 2>/dev/null; python profiles/sass_mix.py $T/s2.csv > $O/sass_mix_bne_state_warp.txt 2>&1
ncu -i $T/dsd.ncu-rep --page source --csv -k regex:dsd_minutes_warp > $T/s3.csv 2>/dev/null; python profiles/sass_mix.py $T/s3.csv > $O/sass_mix_dsd_minutes_warp.txt 2>&1
du -sh gpurun_out
