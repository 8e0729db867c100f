#!/bin/bash
# Development: blocks per SM of the correspondence search grid (persistent warps striding over the queries)
python tools/k1_probe.py 6 || exit 1
for b in auto 4 5 6 8; do
  echo "== K4 blocks/SM $b"
  if [ $b = auto ]; then unset NGICP_K4_BLOCKS_PER_SM; else export NGICP_K4_BLOCKS_PER_SM=$b; fi
  NGICP_BENCH_CFG5_SCANS=0 NGICP_BENCH_CFG4_SCANS=0 timeout 600 python bench.py --steps 40 --warmup 3 2> /dev/null > /tmp/line.json
  python tools/bench_brief.py < /tmp/line.json 2>&1 | head -3 | cut -c1-220
done
