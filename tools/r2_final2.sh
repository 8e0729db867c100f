#!/bin/bash
# Final check after the last edits: GPU tests, smoke, bench line
O=gpurun_out/r02e
mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee $O/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_1gpu.json 2> $O/bench_1gpu.err || { echo "bench failed"; tail -5 $O/bench_1gpu.err; }
python tools/bench_brief.py < $O/bench_1gpu.json 2>&1 | grep "^value\|^{'K1\|^bulk:\|per_rank\|^cpu" | cut -c1-400
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02e/bench_1gpu.json"))
c = d["odom_loop_cfg4"]
print("cfg4 python", round(c["ms_per_scan_median"], 3), "cpp", {k: round(v, 3) for k, v in c["cpp_loop"].items() if k.startswith("ms_")}, "same", c["cpp_loop"]["same_decisions_as_python_loop"], c["oracle"]["same_submap_sets"])
print("cfg5", round(d["multi_sequence_8"]["scans_per_s"]), "e2e", round(d["e2e"]["value"]), round(d["e2e_pageable"]["value"]), "roofline", round(d["roofline"]["frac"], 4), d["roofline"]["traffic"])
PY
