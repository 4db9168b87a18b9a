#!/bin/bash
# one full ncu capture of a chosen kernel (default k_accumulate) at 2^LG points
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LG=${LG:-22}; KERNEL=${KERNEL:-k_accumulate}; SKIP=${SKIP:-2}
SIZES=$LG NOPINT=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s $SKIP -c 1 \
   -f -o gpurun_out/prof_${KERNEL}_$LG python tools/quick_bench.py > gpurun_out/ncu_${KERNEL}_$LG.log 2>&1
tail -4 gpurun_out/ncu_${KERNEL}_$LG.log
ls -la gpurun_out/*.ncu-rep
