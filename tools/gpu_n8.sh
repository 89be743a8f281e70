#!/bin/bash
# the default bench at 8 ranks on one node (weak scaling of the frame pipeline), one short run
TAG=${1:-r02z}
mkdir -p gpurun_out
nproc > gpurun_out/${TAG}_n8_host.txt; python -c "import os; print(len(os.sched_getaffinity(0)))" >> gpurun_out/${TAG}_n8_host.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29708 bench.py --gpus 8 --steps 10 --warmup 3 --cpu-frames 8 > gpurun_out/${TAG}_scale_n8.json 2> gpurun_out/${TAG}_scale_n8.err
python tools/bench_line.py scale_n8 < gpurun_out/${TAG}_scale_n8.json | cut -c1-300
cat gpurun_out/${TAG}_n8_host.txt; tail -3 gpurun_out/${TAG}_scale_n8.err
