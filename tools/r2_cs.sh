#!/bin/bash
# Development: cluster sort inside the filters — parity, then the loop with and without it
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for c in 1 0; do
  echo "== NGICP_SORT_CLUSTER=$c"
  NGICP_SORT_CLUSTER=$c SCANS=150 timeout 300 python tools/odom_profile.py 2>&1 | grep "rep 1" -A1 | cut -c1-300
done
NGICP_BENCH_CFG4_SCANS=0 timeout 600 python bench.py --steps 20 --warmup 3 2> /dev/null > /tmp/line.json; python tools/bench_brief.py < /tmp/line.json 2>&1 | grep "^value\|multi_sequence_8\|prefilter" | cut -c1-400
NGICP_SORT_CLUSTER=0 NGICP_BENCH_CFG4_SCANS=0 timeout 600 python bench.py --steps 20 --warmup 3 2> /dev/null > /tmp/line.json; python tools/bench_brief.py < /tmp/line.json 2>&1 | grep "^value\|multi_sequence_8\|prefilter" | cut -c1-400
