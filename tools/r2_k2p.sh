#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_k2.py -q 2>&1 | tail -2 > gpurun_out/trace.txt
NGICP_K2_TRACE=1 timeout 300 python tools/ab.py k3 2>&1 | grep -E "k2 trace" | tail -2 | head -1 | cut -c1-90 >> gpurun_out/trace.txt
NGICP_K2_TRACE=1 timeout 300 python tools/profile_step.py 8 2>&1 | grep -E "k2 trace" | tail -8 | cut -c1-80 >> gpurun_out/trace.txt
cat gpurun_out/trace.txt
