#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_k2.py -q 2>&1 | tail -5 > gpurun_out/k2_tests.txt; cat gpurun_out/k2_tests.txt
: > gpurun_out/ab_err.log
for m in 4 6; do NGICP_K2_CMAX_MULT=$m timeout 300 python tools/ab.py k3 2>&1 | grep "^K3" >> gpurun_out/ab_err.log; done
cat gpurun_out/ab_err.log
timeout 300 python tools/gpu_diag.py --big 2>&1 | grep -E "^run 2|^untimed run 2|DIAG|FAILED" > gpurun_out/step.txt; cat gpurun_out/step.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"leaf|radix|keys_kernel|bbox|table_build|gather_levels|index_prep|covariance" -c 60 --csv --log-file gpurun_out/k2_launches.csv python tools/profile_step.py 3 > gpurun_out/ncu_list.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/k2_launches.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[-22:]: print(r[4][:60], r[-1], r[-2])
PY
