#!/bin/bash
# where do the GPUs of this box hang (PCIe switch / NUMA node), which CPUs and memory nodes may this
# container use, and what does host<->device bandwidth look like per GPU alone and together?
nvidia-smi topo -m
for g in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader); do
  b=$(echo ${g#0000} | tr A-Z a-z); b="0000${b}"
  echo "gpu $g numa_node=$(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null) local_cpulist=$(cat /sys/bus/pci/devices/$b/local_cpulist 2>/dev/null)"
done
lscpu | grep -i -E "numa|socket|model name|^cpu\(s\)"
echo "cpuset.cpus.effective=$(cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null) mems=$(cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null)"
grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status
which numactl taskset
n=$(nvidia-smi -L | wc -l)
for ((i=0;i<n;i++)); do echo "--- GPU $i alone"; CUDA_VISIBLE_DEVICES=$i tools/pcie_probe | head -3; done
if [ $n -gt 1 ]; then
  echo "--- all $n GPUs at once"
  for ((i=0;i<n;i++)); do (CUDA_VISIBLE_DEVICES=$i tools/pcie_probe 300 sync | sed -n 1,3p | sed "s/^/gpu $i: /") & done; wait
fi
