#!/bin/bash
# Round-2 final validation (GPU box, one GPU): GPU tests, smoke, bench line, reference arm, K1/K4b ncu rows
O=gpurun_out/r02c
mkdir -p $O
SECONDS=0
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $O/pytest_gpu.txt
echo "tests ${SECONDS}s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_1gpu.json 2> $O/bench_1gpu.err || { echo "bench failed"; tail -5 $O/bench_1gpu.err; }
echo "bench ${SECONDS}s"; python tools/bench_brief.py < $O/bench_1gpu.json 2>&1 | cut -c1-1200
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$? ${SECONDS}s"; cut -c1-400 $O/bench_reference.json
timeout 600 ncu --set full --clock-control none -f -k regex:"index_cluster|table_build" --launch-skip 6 -c 2 -o /tmp/k1 python tools/k1_probe.py 6 > /dev/null 2>&1
python tools/summarize_ncu.py /tmp/k1.ncu-rep > $O/ncu_full_K1_kernels.json
KEYFRAMES=8 timeout 600 ncu --set full --clock-control none -f -k regex:"linearize_kernel" -c 2 -o /tmp/k4b python tools/profile_bulk.py > /dev/null 2>&1
python tools/summarize_ncu.py /tmp/k4b.ncu-rep > $O/ncu_full_bulk_K4b.json
python - <<'PY'
import json
for f in ("ncu_full_K1_kernels.json", "ncu_full_bulk_K4b.json"):
    for d in json.load(open("gpurun_out/r02c/" + f)):
        print({k: d[k] for k in ("kernel", "duration_us", "grid", "regs", "warps_active_pct", "issue_active_pct", "dram_traffic_B") if k in d})
PY
echo "done ${SECONDS}s"
