#!/bin/bash
# Development (GPU box): K2 parity, the whole GPU suite, K2 timing (bulk + step)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_k2.py -x -q 2>&1 | tail -15 > gpurun_out/k2_tests.txt; cat gpurun_out/k2_tests.txt
: > gpurun_out/ab_err.log
for chunk in 128 256; do
  NGICP_K2_CHUNK=$chunk timeout 300 python tools/ab.py k3 2>&1 | grep "^K3" >> gpurun_out/ab_err.log
done
cat gpurun_out/ab_err.log
: > gpurun_out/step.txt
for chunk in 128 256; do
  echo "== step, K2_CHUNK=$chunk" >> gpurun_out/step.txt
  NGICP_K2_CHUNK=$chunk timeout 300 python tools/gpu_diag.py --big 2>&1 | grep -E "^run |^untimed|DIAG|FAILED" >> gpurun_out/step.txt
done
cat gpurun_out/step.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
