#!/bin/bash
# Development (GPU box): A/B sweeps of the judged kernels, results in gpurun_out/ab.txt
mkdir -p gpurun_out; : > gpurun_out/ab.txt
for v in 0 1; do NGICP_K3_TMA=$v timeout 300 python tools/ab.py k3 2>>gpurun_out/ab_err.log | tail -1 >> gpurun_out/ab.txt; done
cat gpurun_out/ab.txt
