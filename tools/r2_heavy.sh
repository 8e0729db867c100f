#!/bin/bash
# Development: kLeafHeavy sweep (split threshold of heavy first passes in the leaf search)
for v in 384 512 768 1024; do
  lib=noetic-slam_b200/libngicp_b200_h$v.so; [ $v = 768 ] && lib=noetic-slam_b200/libngicp_b200.so
  echo "== heavy $v"
  NGICP_LIB=$PWD/$lib NGICP_BENCH_CFG5_SCANS=0 NGICP_BENCH_CFG4_SCANS=0 timeout 600 python bench.py --steps 20 --warmup 3 2> /dev/null > /tmp/line.json
  python tools/bench_brief.py < /tmp/line.json 2>&1 | head -3 | cut -c1-200
done
