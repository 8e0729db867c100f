#!/bin/bash
# Development (GPU box): bench.py on 1 GPU
mkdir -p gpurun_out
SECONDS=0
timeout 1200 python bench.py --steps ${STEPS:-20} --warmup 3 2> gpurun_out/bench_err.log > gpurun_out/bench_line.json
echo "bench rc=$? wall ${SECONDS}s"
grep -v "^\[k2" gpurun_out/bench_err.log | tail -8 | cut -c1-400
python tools/bench_brief.py < gpurun_out/bench_line.json > gpurun_out/bench_brief.txt 2>&1
cat gpurun_out/bench_brief.txt
