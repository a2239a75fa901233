#!/bin/bash
mkdir -p gpurun_out
{
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)\|thread"
nvidia-smi topo -m
for d in /sys/bus/pci/devices/*; do c=$(cat $d/class 2>/dev/null); if [[ $c == 0x0302* || $c == 0x0300* ]]; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done
numactl -H 2>/dev/null || echo "no numactl"
cat /proc/self/status | grep -i "cpus_allowed_list\|mems_allowed_list"
nproc
} > gpurun_out/topo.log 2>&1
cat gpurun_out/topo.log
