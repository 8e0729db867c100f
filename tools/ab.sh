#!/bin/bash
# Development (GPU box): quick numbers of the judged bulk kernels under the NGICP_* development switches -> gpurun_out/ab.txt
mkdir -p gpurun_out; : > gpurun_out/ab.txt
timeout 300 python tools/ab.py k3 2>>gpurun_out/ab_err.log | tail -1 >> gpurun_out/ab.txt
timeout 300 python tools/ab.py k4b 2>>gpurun_out/ab_err.log | tail -1 >> gpurun_out/ab.txt
cat gpurun_out/ab.txt
