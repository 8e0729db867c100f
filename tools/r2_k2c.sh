#!/bin/bash
mkdir -p gpurun_out
NGICP_LIB=$PWD/noetic-slam_b200/libngicp_b200_stats.so timeout 300 python tools/k2_debug2.py 2>&1 | grep -v "^\[" > gpurun_out/k2_debug2.txt; cat gpurun_out/k2_debug2.txt
