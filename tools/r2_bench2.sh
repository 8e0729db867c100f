#!/bin/bash
# Development (GPU box, --gpus N): bench.py under torchrun like the driver launches it
N=${N:-2}
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 2> gpurun_out/bench${N}_err.log > gpurun_out/bench${N}_line.json
echo "bench N=$N rc=$? wall ${SECONDS}s"
grep -v "^\[k2\|NCCL\|^W1" gpurun_out/bench${N}_err.log | tail -6 | cut -c1-300
python tools/bench_brief.py < gpurun_out/bench${N}_line.json 2>&1 | grep -v "^cpu\|^odom\|^e2e_cpp\|prefilter" | cut -c1-900
