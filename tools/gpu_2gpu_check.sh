#!/bin/bash
# 2 ranks on one box: the sharded-BA parity tests (library-issued ncclAllReduce) with the current kernels
TAG=${1:-r03h}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_sharding.py -m gpu -x -q > gpurun_out/${TAG}_pytest_sharding.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_sharding.log
tail -4 gpurun_out/${TAG}_pytest_sharding.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tools/sharded_gba_check.py 100 6000 8 > gpurun_out/${TAG}_sharded_gba_2gpu.txt 2>&1
tail -3 gpurun_out/${TAG}_sharded_gba_2gpu.txt
