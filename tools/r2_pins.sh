#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_pins.py tests/test_dropin.py -q 2>&1 | tail -25 > gpurun_out/pins.txt; cat gpurun_out/pins.txt
timeout 600 python bench.py --steps 20 --warmup 3 2> gpurun_out/bench_err.log > gpurun_out/bench_line.json; python tools/bench_brief.py < gpurun_out/bench_line.json 2>&1 | head -3
