#!/bin/bash
for d in 1 0 1 0; do
  NGICP_BENCH_DEDICATE=$d NGICP_BENCH_CFG4_SCANS=0 timeout 600 python bench.py --steps 20 --warmup 3 2> /dev/null > /tmp/line.json
  python - <<PY
import json
d = json.load(open("/tmp/line.json"))
print("dedicate $d: value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "cfg5", round(d["multi_sequence_8"]["scans_per_s"]), "probe4", round(d["bulk"].get("multi_sequence", {}).get("scans_per_s", 0)) if isinstance(d["bulk"].get("multi_sequence"), dict) else "")
PY
done
