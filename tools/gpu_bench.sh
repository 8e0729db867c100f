#!/bin/bash
# Development helper (run on the GPU box through gpurun): bench line -> gpurun_out/bench_line.json, short summary on stdout.
timeout 300 python bench.py --steps ${STEPS:-20} --warmup 3 2> gpurun_out/bench_err.log > gpurun_out/bench_line.json
python tools/bench_brief.py < gpurun_out/bench_line.json > gpurun_out/bench_brief.txt 2>&1
cat gpurun_out/bench_brief.txt
