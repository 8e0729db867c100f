#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_k2.py -q 2>&1 | tail -5 > gpurun_out/k2_tests.txt; cat gpurun_out/k2_tests.txt
NGICP_LIB=$PWD/noetic-slam_b200/libngicp_b200_stats.so timeout 300 python tools/leaf_items.py 2>&1 | grep -v "^\[" > gpurun_out/leaf_items.txt; cat gpurun_out/leaf_items.txt
NGICP_LIB=$PWD/noetic-slam_b200/libngicp_b200_stats.so timeout 300 python tools/leaf_stats.py 2>&1 | grep -v "^\[" | head -3 > gpurun_out/leaf_stats.txt; cat gpurun_out/leaf_stats.txt
: > gpurun_out/ab_err.log
for c2 in 8 4 16; do
  NGICP_K2_CAP2_MULT=$c2 timeout 300 python tools/ab.py k3 2>&1 | grep "^K3" >> gpurun_out/ab_err.log
done
for m in 3 6; do NGICP_K2_CMAX_MULT=$m timeout 300 python tools/ab.py k3 2>&1 | grep "^K3" >> gpurun_out/ab_err.log; done
cat gpurun_out/ab_err.log
: > gpurun_out/step.txt
for c2 in 8 4 16; do
  echo "== step, CAP2_MULT=$c2" >> gpurun_out/step.txt
  NGICP_K2_CAP2_MULT=$c2 timeout 300 python tools/gpu_diag.py --big 2>&1 | grep -E "^run 2|^untimed run 2|DIAG|FAILED" >> gpurun_out/step.txt
done
cat gpurun_out/step.txt
