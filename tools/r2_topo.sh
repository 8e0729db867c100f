#!/bin/bash
# Development: host topology of the GPU box (which logical CPUs are hyper-thread siblings, NUMA nodes, GPU affinity)
lscpu | grep -i "model name\|socket\|core(s)\|thread(s)\|numa\|^CPU(s)" | head -12
cat /sys/devices/system/cpu/cpu0/topology/thread_siblings_list /sys/devices/system/cpu/cpu1/topology/thread_siblings_list 2>/dev/null
python -c "import os; print(sorted(os.sched_getaffinity(0)))"
nvidia-smi topo -m 2>/dev/null | head -14 | cut -c1-200
