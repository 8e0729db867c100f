#!/bin/bash
# development: sweep the knobs of the warp-cooperative search
for c in 32 64 256 1024; do
  echo "== K4_CMAX=$c"; NGICP_K4_CMAX=$c python tools/gpu_diag.py --big 2>&1 | grep -E "^run 2" | sed -E "s/.*'knn_ms': ([0-9.]+).*'correspond_ms': ([0-9.]+).*/knn \1 corr3 \2/"
done
for l in 1 4; do
  echo "== K4_LPQ=$l"; NGICP_K4_LPQ=$l python tools/gpu_diag.py --big 2>&1 | grep -E "^run 2" | sed -E "s/.*'knn_ms': ([0-9.]+).*'correspond_ms': ([0-9.]+).*/knn \1 corr3 \2/"
done
for m in 2 4 8; do
  echo "== K2_CMAX_MULT=$m"; NGICP_K2_CMAX_MULT=$m python tools/gpu_diag.py --big 2>&1 | grep -E "^run 2" | sed -E "s/.*'knn_ms': ([0-9.]+).*'correspond_ms': ([0-9.]+).*/knn \1 corr3 \2/"
done
