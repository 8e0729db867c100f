#!/bin/bash
# development: sweep the knobs of the warp-cooperative search
for l in 1 4; do for c in 32 64 128; do
  echo "== K4_LPQ=$l K4_CMAX=$c"; NGICP_K4_LPQ=$l NGICP_K4_CMAX=$c python tools/gpu_diag.py --big 2>&1 | grep -E "^run 2" | sed -E "s/.*'knn_ms': ([0-9.]+).*'linearize_ms': ([0-9.]+).*/knn \1 lin3 \2/"
done; done
for l in 1 4; do for m in 2 4; do
  echo "== K2_LPQ=$l K2_CMAX_MULT=$m"; NGICP_K2_LPQ=$l NGICP_K2_CMAX_MULT=$m python tools/gpu_diag.py --big 2>&1 | grep -E "^run 2|target cov wall|bad|dT" | sed -E "s/.*'knn_ms': ([0-9.]+).*'linearize_ms': ([0-9.]+).*/knn \1 lin3 \2/"
done; done
