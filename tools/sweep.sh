#!/bin/bash
# development: sweep the group-capacity knobs of the warp-cooperative search
for c in 32 64 128 256 512 1024; do
  echo "== NGICP_K4_CMAX=$c"; NGICP_K4_CMAX=$c python tools/gpu_diag.py --big 2>&1 | grep -E "^run 2" | sed -E "s/.*'knn_ms': ([0-9.]+).*'linearize_ms': ([0-9.]+).*/knn \1 lin3 \2/"
done
for m in 2 3 4 6; do
  echo "== NGICP_K2_CMAX_MULT=$m"; NGICP_K2_CMAX_MULT=$m python tools/gpu_diag.py --big 2>&1 | grep -E "^run 2|target cov wall" | sed -E "s/.*'knn_ms': ([0-9.]+).*'linearize_ms': ([0-9.]+).*/knn \1 lin3 \2/"
done
