#!/bin/bash
mkdir -p gpurun_out
NGICP_LIB=$PWD/noetic-slam_b200/libngicp_b200_stats.so timeout 300 python tools/leaf_items.py 2>&1 | grep -v "^\[" > gpurun_out/leaf_items.txt; cat gpurun_out/leaf_items.txt
