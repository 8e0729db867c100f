#!/bin/bash
# Round-2 profile capture (GPU box, one GPU): bench line first (no profiler), then the ncu launch list of the bench command
# (odom-loop legs off: the list is of the headline step, the e2e arms and the cfg-3 bulk build), then ncu --set full of the
# step kernels (profiler range = two steps) and of the bulk kernels. Only summaries are kept (the .ncu-rep files stay in /tmp).
O=gpurun_out/r02
mkdir -p $O
SECONDS=0
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_1gpu.json 2> $O/bench_1gpu.err || { echo "bench failed"; tail -5 $O/bench_1gpu.err; exit 1; }
echo "bench ok ${SECONDS}s"; python tools/bench_brief.py < $O/bench_1gpu.json 2>&1 | head -4
NGICP_BENCH_CFG5_SCANS=0 NGICP_BENCH_CFG4_SCANS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/bench_launches.csv \
  python bench.py --steps 2 --warmup 3 > /tmp/bench_under_ncu.json 2> /tmp/bench_under_ncu.err
echo "launch list rc=$? ${SECONDS}s"; python tools/summarize_launches.py $O/bench_launches.csv > $O/bench_launches_summary.txt; head -14 $O/bench_launches_summary.txt
python tools/profile_step.py 2 > /dev/null 2>&1 || { echo "profile_step failed"; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o /tmp/step python tools/profile_step.py 2 > $O/step_ncu.log 2>&1
echo "step full rc=$? ${SECONDS}s"; python tools/summarize_ncu.py /tmp/step.ncu-rep > $O/ncu_full_step_kernels.json
KEYFRAMES=256 timeout 900 ncu --set full --clock-control none --import-source on -f -k regex:"covariance_kernel|linearize_kernel|knn_leaf_kernel|leaf_items_kernel" -c 10 -o /tmp/bulk python tools/profile_bulk.py > $O/bulk_ncu.log 2>&1
echo "bulk full rc=$? ${SECONDS}s"; python tools/summarize_ncu.py /tmp/bulk.ncu-rep > $O/ncu_full_bulk_kernels.json
python tools/ncu_lines.py /tmp/bulk.ncu-rep covariance_kernel noetic-slam_b200/csrc/build/covariance.o > $O/k3_lines.txt 2>&1
ls -la $O; du -sh gpurun_out
