#!/bin/bash
# Round-2 closing evidence, run on the GPU box: bash profiles/collect_r2_final.sh   (outputs under gpurun_out/r2f/)
set -u
O=gpurun_out/r2f; mkdir -p $O/sweep
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,launch__grid_size,launch__block_size"
python -m pytest tests -x -q -m gpu > $O/pytest_gpu.txt 2>&1; tail -2 $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
python bench.py > $O/bench_default.out 2> $O/bench_default.err; tail -1 $O/bench_default.out > $O/bench_default.json
python bench.py --impl reference > $O/bench_reference.out 2> $O/bench_reference.err; tail -1 $O/bench_reference.out > $O/bench_reference.json
python bench.py --fft tc --no-cpu 2> $O/bench_tc.err | tail -1 > $O/bench_tc.json
python bench.py --config 1 2> $O/bench_config1.err | tail -1 > $O/bench_config1.json
python bench.py --config 4 2> $O/bench_config4.err | tail -1 > $O/bench_config4.json
for w in bne roe dsd; do python bench.py --workload $w --steps 3 --warmup 3 2> $O/nb_$w.err | tail -1 > $O/bench_$w.json; done
echo "benches done"
python profiles/sweep_bench.py 3600 > $O/sweep_features_1h_r2.jsonl 2> $O/sweep.err
for g in "256 128" "256 64" "512 256" "512 128" "1024 512" "1024 256" "2048 1024" "2048 512" "4096 2048" "4096 1024"; do
  set -- $g
  timeout 200 ncu --metrics $M --clock-control none -k regex:"stft" -s 7 -c 1 --csv --log-file $O/sweep/stft_$1_$2_f64.csv python profiles/sweep_bench.py 3600 $1 $2 f64 > /dev/null 2>&1
done
timeout 200 ncu --metrics $M --clock-control none -k regex:"tcdft" -s 3 -c 1 --csv --log-file $O/sweep/tcdft_256_128.csv python profiles/sweep_bench.py 3600 256 128 tc > /dev/null 2>&1
echo "sweep ncu done"
du -sh gpurun_out; ls -la $O | tail -30
