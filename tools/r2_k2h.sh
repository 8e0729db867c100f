#!/bin/bash
mkdir -p gpurun_out
NGICP_K2_TRACE=1 timeout 300 python tools/ab.py k3 2>&1 | grep -E "^K3|k2 trace" > gpurun_out/trace.txt
NGICP_K2_TRACE=1 timeout 300 python tools/gpu_diag.py --big 2>&1 | grep -E "^run 2|k2 trace" | tail -8 >> gpurun_out/trace.txt
cat gpurun_out/trace.txt
