#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_k2.py -q 2>&1 | tail -3 > gpurun_out/trace.txt
NGICP_K2_TRACE=1 timeout 300 python tools/ab.py k3 2>&1 | grep -E "^K3|k2 trace" | tail -3 | head -1 >> gpurun_out/trace.txt
NGICP_K2_TRACE=1 timeout 300 python tools/gpu_diag.py --big 2>&1 | grep -E "k2 trace" | tail -5 | head -2 >> gpurun_out/trace.txt
NGICP_K2_TRACE=1 timeout 300 python tools/profile_step.py 8 2>&1 | grep -E "k2 trace" | tail -8 >> gpurun_out/trace.txt
cat gpurun_out/trace.txt
NGICP_LIB=$PWD/noetic-slam_b200/libngicp_b200_stats.so timeout 300 python tools/leaf_items.py bench 2>&1 | grep -v "^\[" | head -12 > gpurun_out/leaf_items.txt; cat gpurun_out/leaf_items.txt
