cd "$(dirname "$0")/.."
: > gpurun_out/r02c_ncu_full_sort.csv
for k in k_coarse_count k_partition_coarse k_fine_smem; do
  SIZES=24 NOPINT=1 timeout 300 ncu --set full --clock-control none -k regex:^$k -s 1 -c 1 -f -o gpurun_out/s_$k python tools/quick_bench.py > /dev/null 2>&1
  python tools/ncu_summary.py gpurun_out/s_$k.ncu-rep >> gpurun_out/r02c_ncu_full_sort.csv
  rm -f gpurun_out/s_$k.ncu-rep
done
