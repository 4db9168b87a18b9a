#!/bin/bash
# Development aid: host topology of the GPU box (NUMA nodes, which CPUs are local to which GPU).
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)\|thread"
nvidia-smi topo -m 2>&1 | head -30
for d in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader); do
  b=$(echo $d | tr 'A-Z' 'a-z' | sed 's/^0000//')
  echo "$d numa_node=$(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null) local_cpulist=$(cat /sys/bus/pci/devices/$b/local_cpulist 2>/dev/null)"
done
nproc; cat /proc/self/status | grep -i "cpus_allowed_list\|mems_allowed_list"
python - <<'PY'
import os
print("affinity", sorted(os.sched_getaffinity(0)))
PY
