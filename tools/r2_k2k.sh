#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:knn_leaf_kernel -s 2 -c 1 -f -o gpurun_out/r02c_k2scan python tools/profile_step.py 3 > gpurun_out/ncu_k2scan.log 2>&1
ls -la gpurun_out/r02c_k2scan.ncu-rep
