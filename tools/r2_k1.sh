#!/bin/bash
# Development (GPU box): K1 cluster build — parity first, then the step times with and without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
for c in 1 0; do
  echo "== NGICP_K1_CLUSTER=$c"
  NGICP_K1_CLUSTER=$c NGICP_BENCH_CFG5_SCANS=0 NGICP_BENCH_CFG4_SCANS=0 timeout 600 python bench.py --steps 20 --warmup 3 2> gpurun_out/k1_err_$c.log > gpurun_out/k1_line_$c.json
  python tools/bench_brief.py < gpurun_out/k1_line_$c.json 2>&1 | head -3
done
