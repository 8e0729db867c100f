#!/bin/bash
# Round-2 profile capture (GPU box, one GPU): bench line first (no profiler), then the ncu launch list of the bench command,
# then ncu --set full of the step kernels and of the bulk kernels. Summaries land in gpurun_out/r02/.
O=gpurun_out/r02
mkdir -p $O
SECONDS=0
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_1gpu.json 2> $O/bench_1gpu.err || { echo "bench failed"; tail -5 $O/bench_1gpu.err; exit 1; }
echo "bench ok ${SECONDS}s"; python tools/bench_brief.py < $O/bench_1gpu.json 2>&1 | head -4
# launch list of the same command (the odom-loop legs bounded so that the list stays readable; the headline step, e2e, cfg-3 bulk are all in)
NGICP_BENCH_CFG5_SCANS=0 NGICP_BENCH_CFG4_SCANS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/bench_launches.csv \
  python bench.py --steps 2 --warmup 3 > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
echo "launch list rc=$? ${SECONDS}s"; python tools/summarize_launches.py $O/bench_launches.csv > $O/bench_launches_summary.txt; head -12 $O/bench_launches_summary.txt
python tools/profile_step.py 2 > /dev/null 2>&1 || { echo "profile_step failed"; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -f -o $O/step python tools/profile_step.py 2 > $O/step_ncu.log 2>&1
echo "step full rc=$? ${SECONDS}s"; python tools/summarize_ncu.py $O/step.ncu-rep > $O/ncu_full_step_kernels.json
KEYFRAMES=256 python tools/profile_bulk.py > $O/bulk_plain.txt 2>&1 || { echo "profile_bulk failed"; tail -3 $O/bulk_plain.txt; exit 1; }
tail -1 $O/bulk_plain.txt
KEYFRAMES=256 timeout 1500 ncu --set full --clock-control none --import-source on -f -k regex:"covariance_kernel|linearize_kernel|knn_leaf_kernel|leaf_items_kernel|correspond" -c 24 -o $O/bulk python tools/profile_bulk.py > $O/bulk_ncu.log 2>&1
echo "bulk full rc=$? ${SECONDS}s"; python tools/summarize_ncu.py $O/bulk.ncu-rep > $O/ncu_full_bulk_kernels.json
ls -la $O
