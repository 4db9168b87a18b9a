cd "$(dirname "$0")/.."
for k in k_fine_local k_partition_fine k_partition_coarse k_coarse_count; do
  v2=1; [ $k = k_partition_fine ] && v2=""
  env ${v2:+PORLA_SORT_V2=1} SIZES=24 NOPINT=1 timeout 300 ncu --set full --clock-control none -k regex:^$k\$ -s 1 -c 1 -f -o gpurun_out/v2_$k python tools/quick_bench.py > /dev/null 2>&1
  python tools/ncu_summary.py gpurun_out/v2_$k.ncu-rep >> gpurun_out/v2_sort_ncu.csv
  rm -f gpurun_out/v2_$k.ncu-rep
done
