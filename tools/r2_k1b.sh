#!/bin/bash
mkdir -p gpurun_out
python tools/k1_probe.py 8 || exit 1
timeout 600 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:"index_cluster|table_build" --launch-skip 6 -c 2 -o /tmp/k1 python tools/k1_probe.py 6 > /tmp/k1_ncu.log 2>&1
echo "ncu rc=$?"
python tools/summarize_ncu.py /tmp/k1.ncu-rep > gpurun_out/k1_ncu.json
python - <<'PY'
import json
for d in json.load(open("gpurun_out/k1_ncu.json")):
    print({k: d[k] for k in ("kernel", "duration_us", "grid", "regs", "warps_active_pct", "issue_active_pct", "inst_executed", "dram_traffic_B") if k in d})
PY
python tools/ncu_lines.py /tmp/k1.ncu-rep index_cluster_kernelILi4 noetic-slam_b200/csrc/build/index.o 2>&1 | head -${LINES_N:-24}
