#!/bin/bash
# essential-graph engine on the GPU: parity tests, then timing of one 400-keyframe graph
TAG=${1:-r02w}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_posegraph.py -m gpu -x -q > gpurun_out/${TAG}_pytest_pg.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_pg.log
tail -30 gpurun_out/${TAG}_pytest_pg.log
timeout 200 python tools/posegraph_bench.py > gpurun_out/${TAG}_posegraph_bench.txt 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_posegraph_bench.txt
cat gpurun_out/${TAG}_posegraph_bench.txt
timeout 300 python -m pytest tests/test_gpu_shims.py -m gpu -x -q > gpurun_out/${TAG}_pytest_shims.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_shims.log
tail -15 gpurun_out/${TAG}_pytest_shims.log
