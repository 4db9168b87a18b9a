#!/bin/bash
# Development aid: stage timings under several tunables in one GPU call.  Usage: tools/exp.sh "ENV1=a ENV2=b" "ENV1=c" ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/exp.txt
for cfg in "$@"; do
  echo "### $cfg" | tee -a gpurun_out/exp.txt
  env $cfg SIZES=${SIZES:-16,20,24} timeout 300 python tools/stage_times.py 2>&1 | tee -a gpurun_out/exp.txt
done
