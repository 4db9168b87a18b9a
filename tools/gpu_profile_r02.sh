#!/bin/bash
# Round-2 profile set: launch list of the bench command, full ncu captures of k_accumulate at 2^20 / 2^24, sanitizer runs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:porla -c 1500 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sweep "" > gpurun_out/r02_bench_under_ncu.log 2>&1
tail -2 gpurun_out/r02_bench_under_ncu.log | cut -c1-300
for lg in 20 24; do
  SIZES=$lg NOPINT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 2 -c 1 \
     -f -o gpurun_out/r02_prof_k_accumulate_$lg python tools/quick_bench.py > gpurun_out/r02_ncu_acc_$lg.log 2>&1
  tail -2 gpurun_out/r02_ncu_acc_$lg.log
done
SAN_BIG=1 timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py > gpurun_out/r02_compute_sanitizer_memcheck.log 2>&1
tail -3 gpurun_out/r02_compute_sanitizer_memcheck.log
SAN_BIG=0 timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py > gpurun_out/r02_compute_sanitizer_racecheck.log 2>&1
tail -3 gpurun_out/r02_compute_sanitizer_racecheck.log
