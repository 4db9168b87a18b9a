#!/bin/bash
# per-launch device times (ncu, cold-cache/serialised: compare shares)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lg in ${SIZES_LIST:-16 20 24}; do
  SIZES=$lg NOPINT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv \
     --log-file gpurun_out/launches_$lg.csv python tools/quick_bench.py > gpurun_out/qb_$lg.log 2>&1
  tail -3 gpurun_out/qb_$lg.log
done
