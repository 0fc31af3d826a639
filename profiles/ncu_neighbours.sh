#!/bin/bash
# ncu --set full of the neighbour engines' main kernels; the reports stay on the box (too large to bring back):
# their per-kernel summaries and per-opcode stall mixes come back as text.  bash profiles/ncu_neighbours.sh
set -u
O=gpurun_out/r2nb; mkdir -p $O; T=/tmp/ncu_nb; mkdir -p $T
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"roe_filter_wave|roe_frame|roe_part" -c 3 -o $T/roe python bench.py --workload roe --steps 1 --warmup 0 --no-cpu > /dev/null 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"bne_filter_wave|bne_state_warp|bne_fft512" -c 3 -o $T/bne python bench.py --workload bne --steps 1 --warmup 0 --no-cpu > /dev/null 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"dsd_" -c 2 -o $T/dsd python bench.py --workload dsd --steps 1 --warmup 0 --no-cpu > /dev/null 2>&1
for w in roe bne dsd; do
  python profiles/ncu_summary.py $T/$w.ncu-rep > $O/ncu_summary_$w.txt 2>&1
done
ncu -i $T/roe.ncu-rep --page source --csv -k regex:roe_filter_wave > $T/roe_src.csv 2>/dev/null; python profiles/sass_mix.py $T/roe_src.csv > $O/sass_mix_roe_filter_wave.txt 2>&1
ncu -i $T/bne.ncu-rep --page source --csv -k regex:bne_filter_wave > $T/bnef_src.csv 2>/dev/null; python profiles/sass_mix.py $T/bnef_src.csv > $O/sass_mix_bne_filter_wave.txt 2>&1
ncu -i $T/bne.ncu-rep --page source --csv -k regex:bne_state_warp > $T/bnes_src.csv 2>/dev/null; python profiles/sass_mix.py $T/bnes_src.csv > $O/sass_mix_bne_state_warp.txt 2>&1
head -c 3000 $O/ncu_summary_roe.txt
