#!/bin/bash
# Development (GPU box): K2 parity, timing, then ncu --set full of the leaf search (bulk launch and single-scan launch).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_k2.py -x -q 2>&1 | tail -5 > gpurun_out/k2_tests.txt; cat gpurun_out/k2_tests.txt
: > gpurun_out/ab.txt
for chunk in 256 128; do
  NGICP_K2_CHUNK=$chunk timeout 300 python tools/ab.py k3 2>>gpurun_out/ab_err.log | tail -1 >> gpurun_out/ab.txt
done
cat gpurun_out/ab.txt
: > gpurun_out/step.txt
for chunk in 256 128; do
  echo "== step, K2_CHUNK=$chunk" >> gpurun_out/step.txt
  NGICP_K2_CHUNK=$chunk timeout 300 python tools/gpu_diag.py --big 2>&1 | grep -E "^run |^untimed|DIAG|FAILED" >> gpurun_out/step.txt
done
cat gpurun_out/step.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:knn_leaf_kernel -s 1 -c 1 -f -o gpurun_out/r02_k2bulk python tools/ab.py k3 > gpurun_out/ncu_k2bulk.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"knn_leaf_kernel|leaf_items" -s 2 -c 2 -f -o gpurun_out/r02_k2scan python tools/profile_step.py 3 > gpurun_out/ncu_k2scan.log 2>&1
ls -la gpurun_out/*.ncu-rep
