#!/bin/bash
# SCALE-style evidence on ONE box (gpurun --gpus 4): the default bench (configs[1]) at 1 / 2 / 4 GPUs, the sharded-BA parity
# checks at 2 ranks (library-issued ncclAllReduce, skewed abort) and the 2-GPU pytest.  Outputs under gpurun_out/<tag>_*
TAG=${1:-r}
NG=${2:-4}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "${@:2}"; }
timeout 200 python -m pytest tests/test_sharding.py -m gpu -q > gpurun_out/${TAG}_pytest_sharding.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_sharding.log
timeout 200 bash -c "$(declare -f run); run 2 tools/sharded_lba_check.py" > gpurun_out/${TAG}_sharded_lba_2gpu.txt 2>&1; tail -2 gpurun_out/${TAG}_sharded_lba_2gpu.txt
timeout 300 bash -c "$(declare -f run); run 2 tools/sharded_gba_check.py 100 6000 8" > gpurun_out/${TAG}_sharded_gba_2gpu.txt 2>&1; tail -2 gpurun_out/${TAG}_sharded_gba_2gpu.txt
for n in 1 2 4 8; do
  [ $n -gt $NG ] && continue
  if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n))"; fi
  timeout 300 $L bench.py --gpus $n --steps 10 --warmup 3 --cpu-frames 4 > gpurun_out/${TAG}_scale_n$n.json 2> gpurun_out/${TAG}_scale_n$n.err
  python tools/bench_line.py scale_n$n < gpurun_out/${TAG}_scale_n$n.json | cut -c1-200
done
