#!/bin/bash
# Round-2c profile set (final tree): launch list of the bench command, full ncu captures of k_accumulate (traffic figure),
# of the quad kernels (k_reduce_scan at 2^20 / 2^16, k_butterfly_quad at n = 1024, k_lut_sum of a 128-term commitment),
# sanitizer runs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 1500 --csv --log-file gpurun_out/r02c_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sweep "" > gpurun_out/r02c_bench_under_ncu.log 2>&1
tail -2 gpurun_out/r02c_bench_under_ncu.log | cut -c1-200
cap() {  # name kernel skip command...
  local name=$1 kern=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 -f -o gpurun_out/r02c_prof_$name "$@" > gpurun_out/r02c_ncu_$name.log 2>&1
  tail -1 gpurun_out/r02c_ncu_$name.log | cut -c1-200
  python tools/ncu_summary.py gpurun_out/r02c_prof_$name.ncu-rep >> gpurun_out/r02c_ncu_full.csv
  rm -f gpurun_out/r02c_prof_$name.ncu-rep      # (five reports exceed what gpurun copies back; the summary is what profiles/ keeps)
}
: > gpurun_out/r02c_ncu_full.csv
SIZES=20 NOPINT=1 cap k_accumulate_20 k_accumulate 2 python tools/quick_bench.py
SIZES=20 NOPINT=1 cap k_reduce_scan_20 k_reduce_scan 2 python tools/quick_bench.py
SIZES=16 NOPINT=1 cap k_reduce_scan_16 k_reduce_scan 2 python tools/quick_bench.py
NS=1024 cap k_butterfly_quad k_butterfly_quad 3 python tools/butterfly_quad_times.py
cap k_lut_sum_digest k_lut_sum 30 python tools/small_latency.py
SAN_BIG=1 timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py > gpurun_out/r02c_compute_sanitizer_memcheck.log 2>&1
tail -3 gpurun_out/r02c_compute_sanitizer_memcheck.log
SAN_BIG=0 timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py > gpurun_out/r02c_compute_sanitizer_racecheck.log 2>&1
tail -3 gpurun_out/r02c_compute_sanitizer_racecheck.log
SAN_BIG=0 timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_run.py > gpurun_out/r02c_compute_sanitizer_synccheck.log 2>&1
tail -3 gpurun_out/r02c_compute_sanitizer_synccheck.log
ls -la gpurun_out/r02c_*
