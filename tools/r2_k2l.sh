#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/trace.txt
for tma in 1 0; do
echo "== K2_TMA=$tma" >> gpurun_out/trace.txt
NGICP_K2_TMA=$tma timeout 300 python -m pytest tests/test_gpu_k2.py -q 2>&1 | tail -2 >> gpurun_out/trace.txt
NGICP_K2_TMA=$tma NGICP_K2_TRACE=1 timeout 300 python tools/ab.py k3 2>&1 | grep -E "^K3|k2 trace" | tail -3 | head -1 >> gpurun_out/trace.txt
NGICP_K2_TMA=$tma NGICP_K2_TRACE=1 timeout 300 python tools/gpu_diag.py --big 2>&1 | grep -E "k2 trace" | tail -5 | head -2 >> gpurun_out/trace.txt
NGICP_K2_TMA=$tma NGICP_K2_TRACE=1 timeout 300 python tools/profile_step.py 4 2>&1 | grep -E "k2 trace" | tail -2 >> gpurun_out/trace.txt
done
cat gpurun_out/trace.txt
