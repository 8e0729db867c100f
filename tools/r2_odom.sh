#!/bin/bash
# Development (GPU box): the C++ odom loop against the Python loop, then the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_odom.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/odom_tests.txt
cat gpurun_out/odom_tests.txt
SECONDS=0
timeout 1200 python bench.py --steps 20 --warmup 3 2> gpurun_out/bench1_err.log > gpurun_out/bench1_line.json
echo "bench rc=$? wall ${SECONDS}s"
tail -5 gpurun_out/bench1_err.log | cut -c1-300
python tools/bench_brief.py < gpurun_out/bench1_line.json 2>&1 | cut -c1-1500
