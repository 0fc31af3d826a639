#!/bin/bash
# Round-2 evidence, run on the GPU box: bash profiles/collect_r2.sh   (text / csv outputs under gpurun_out/r2/; the
# .ncu-rep files are summarised on the box and deleted: gpurun brings back at most 64 MiB)
set -u
O=gpurun_out/r2; mkdir -p $O/sweep
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,launch__grid_size,launch__block_size"
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
# 1. launch list of the bench command itself (shares of the step)
$B > $O/bench_plain.json 2> $O/bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/launches_bench_r2.csv $B > $O/ncu_bench.log 2>&1
echo "launch list rc=$?"
# 2. every kernel once, full set, on a workload beyond L2 (16 x 600 s: planes of 238 MB each)
export APT_SEGMENTS=1
python profiles/run_small.py 16 600 > $O/small_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k regex:"kernel" -c 26 -o /tmp/full_16x600 python profiles/run_small.py 16 600 > $O/ncu_full.log 2>&1
echo "full capture rc=$?"
ncu -i /tmp/full_16x600.ncu-rep --page raw --csv > $O/ncu_full_16x600_raw.csv 2>/dev/null
python profiles/ncu_summary.py /tmp/full_16x600.ncu-rep > $O/ncu_summary_r2.txt 2>&1
python profiles/make_traffic.py /tmp/full_16x600.ncu-rep $((16 * 52322)) $O/kernel_traffic_r2.json "profiles/run_small.py 16 600 (16 clips x 600 s, planes of 238 MB: beyond the 126 MB L2), APT_SEGMENTS=1" > $O/traffic.log 2>&1
# 3. the tensor-core front end
python profiles/run_small.py 16 600 tc > $O/small_tc_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none -k regex:tcdft -c 1 -o /tmp/full_tcdft python profiles/run_small.py 16 600 tc > $O/ncu_tc.log 2>&1
python profiles/ncu_summary.py /tmp/full_tcdft.ncu-rep > $O/ncu_summary_tcdft_r2.txt 2>&1
ncu -i /tmp/full_tcdft.ncu-rep --page raw --csv > $O/ncu_tcdft_raw.csv 2>/dev/null
unset APT_SEGMENTS
echo "tc capture rc=$?"
# 4. frame-size sweep (BASELINE configs[4]): timings, then one ncu pass per geometry (float64 FFT; the GEMM variant at 256/128)
python profiles/sweep_bench.py 3600 > $O/sweep_features_1h_r2.jsonl 2> $O/sweep.err
echo "sweep rc=$?"
for g in "256 128" "256 64" "512 256" "512 128" "1024 512" "1024 256" "2048 1024" "2048 512" "4096 2048" "4096 1024"; do
  set -- $g
  python profiles/sweep_bench.py 3600 $1 $2 f64 > /dev/null 2>&1 && \
  timeout 200 ncu --metrics $M --clock-control none -k regex:"stft" -s 7 -c 1 --csv --log-file $O/sweep/stft_$1_$2_f64.csv python profiles/sweep_bench.py 3600 $1 $2 f64 > /dev/null 2>&1
done
python profiles/sweep_bench.py 3600 256 128 tc > /dev/null 2>&1 && \
timeout 200 ncu --metrics $M --clock-control none -k regex:"tcdft" -s 3 -c 1 --csv --log-file $O/sweep/tcdft_256_128.csv python profiles/sweep_bench.py 3600 256 128 tc > /dev/null 2>&1
echo "sweep ncu done"
rm -f /tmp/*.ncu-rep
du -sh gpurun_out; ls -la $O $O/sweep | tail -40
