#!/bin/bash
# Development: leaf rule of K2 (cmax = mult * k points per leaf; parent <= cap2 * cmax) for single scans
timeout 800 python -m pytest tests/test_gpu_cluster.py -q -m gpu 2>&1 | tail -6
python tools/k2_probe.py > /dev/null 2>&1
for cm in 3 4 6 8; do for c2 in 4 8 16; do
  echo -n "cmax_mult $cm cap2_mult $c2: "
  NGICP_K2_CMAX_MULT=$cm NGICP_K2_CAP2_MULT=$c2 python tools/k2_probe.py 2>&1 | tail -1
done; done
