#!/bin/bash
# Round-1 (e) evidence run: sanitizer over every kernel, ncu launch list of the bench command, full captures of the
# dominant kernel at 2^20 / 2^24 and of the two one-launch small-MSM kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SAN_BIG=1 timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py > gpurun_out/san_memcheck.log 2>&1; tail -3 gpurun_out/san_memcheck.log
SAN_BIG=0 timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py > gpurun_out/san_racecheck.log 2>&1; tail -3 gpurun_out/san_racecheck.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 600 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --sweep "" > gpurun_out/bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/bench_under_ncu.log
for lg in 20 24; do
  SIZES=$lg NOPINT=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 2 -c 1 \
     -f -o gpurun_out/prof_k_accumulate_$lg python tools/quick_bench.py > gpurun_out/ncu_acc_$lg.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lut_sum -s 30 -c 1 -f -o gpurun_out/prof_k_lut_sum \
    python tools/small_latency.py > gpurun_out/ncu_lut.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_small_bits -s 30 -c 1 -f -o gpurun_out/prof_k_small_bits \
    python tools/small_latency.py > gpurun_out/ncu_bits.log 2>&1
ls -la gpurun_out/*.ncu-rep
