#!/bin/bash
# Development (GPU box): K2 parity tests, then old-vs-new K2 timing on the bulk batch and the single-scan step.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_k2.py -x -q > gpurun_out/k2_tests.txt 2>&1; echo "k2 tests rc=$?" | tee -a gpurun_out/k2_tests.txt
tail -15 gpurun_out/k2_tests.txt
: > gpurun_out/ab.txt
for leaf in 0 1; do
  for chunk in 256 128; do
    [ $leaf = 0 ] && [ $chunk = 128 ] && continue
    NGICP_K2_LEAF=$leaf NGICP_K2_CHUNK=$chunk timeout 300 python tools/ab.py k3 2>>gpurun_out/ab_err.log | tail -1 >> gpurun_out/ab.txt
  done
done
for m in 2 8; do NGICP_K2_CMAX_MULT=$m timeout 300 python tools/ab.py k3 2>>gpurun_out/ab_err.log | tail -1 >> gpurun_out/ab.txt; done
cat gpurun_out/ab.txt
for leaf in 0 1; do
  echo "== step, K2_LEAF=$leaf" >> gpurun_out/step.txt
  NGICP_K2_LEAF=$leaf timeout 300 python tools/gpu_diag.py --big 2>&1 | grep -E "^run |^untimed|DIAG|FAILED|rows exact|max abs err" >> gpurun_out/step.txt
done
cat gpurun_out/step.txt
