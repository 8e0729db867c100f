#!/bin/bash
# Development: K3 with fp32 neighbour differences (-DNGICP_K3_F32_DIFF) against the default build
for v in f32 def; do
  lib=$PWD/noetic-slam_b200/libngicp_b200_f32.so; [ $v = def ] && lib=$PWD/noetic-slam_b200/libngicp_b200.so
  echo "== $v"
  NGICP_LIB=$lib timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pins.py -x -q -m gpu -k "covarian or degenerate or regular or align or odom or pose" 2>&1 | tail -3
  NGICP_LIB=$lib NGICP_BENCH_CFG5_SCANS=0 NGICP_BENCH_CFG4_SCANS=0 timeout 600 python bench.py --steps 20 --warmup 3 2> /dev/null > /tmp/line.json
  python tools/bench_brief.py < /tmp/line.json 2>&1 | head -3 | cut -c1-220
done
