#!/bin/bash
# Development (GPU box): A/B sweeps of the judged kernels, results in gpurun_out/ab.txt
mkdir -p gpurun_out; : > gpurun_out/ab.txt
for v in 0 1 2 3; do NGICP_K3_ICVT=$v timeout 300 python tools/ab.py k3 2>>gpurun_out/ab_err.log | tail -1 >> gpurun_out/ab.txt; done
for pf in 0 1; do for m in 4 8 12; do NGICP_K4B_PF=$pf NGICP_K4B_MULT=$m timeout 300 python tools/ab.py k4b 2>>gpurun_out/ab_err.log | tail -1 >> gpurun_out/ab.txt; done; done
cat gpurun_out/ab.txt
