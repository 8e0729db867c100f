#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/trace.txt
for lib in libngicp_b200.so libngicp_b200_b5.so libngicp_b200_b4.so; do
echo "== $lib" >> gpurun_out/trace.txt
export NGICP_LIB=$PWD/noetic-slam_b200/$lib
timeout 300 python -m pytest tests/test_gpu_k2.py -q 2>&1 | tail -1 >> gpurun_out/trace.txt
NGICP_K2_TRACE=1 timeout 300 python tools/ab.py k3 2>&1 | grep -E "k2 trace" | tail -2 | head -1 >> gpurun_out/trace.txt
NGICP_K2_TRACE=1 timeout 300 python tools/profile_step.py 8 2>&1 | grep -E "k2 trace" | tail -8 | awk '{s+=$9; n++} END {print "mean search ms over", n, "scans:", s/n}' >> gpurun_out/trace.txt
done
cat gpurun_out/trace.txt
