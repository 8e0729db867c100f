#!/bin/bash
# Development (GPU box): knob sweeps of the bulk k-NN, results in gpurun_out/ab.txt
mkdir -p gpurun_out; : > gpurun_out/ab.txt
for l in 1 4; do for m in 2 4 8 16; do NGICP_K2_LPQ=$l NGICP_K2_CMAX_MULT=$m timeout 300 python tools/ab.py k3 2>>gpurun_out/ab_err.log | tail -1 >> gpurun_out/ab.txt; done; done
cat gpurun_out/ab.txt
