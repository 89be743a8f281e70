#!/bin/bash
# Short gpurun call: parity of the SURVEY 8(f) kernels + the guided searches they touch, then a 3-step bench.
TAG=${1:-v}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_next.py tests/test_gpu_sbp.py tests/test_gpu_imu.py -q -m gpu > gpurun_out/${TAG}_pytest_next.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_next.log
tail -25 gpurun_out/${TAG}_pytest_next.log
timeout 240 python bench.py --steps 3 --warmup 3 --cpu-frames 8 > gpurun_out/${TAG}_bench_short.json 2> gpurun_out/${TAG}_bench_short.err; echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_bench_short.err
python -c "
import json,sys
d = json.loads(open('gpurun_out/${TAG}_bench_short.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['stage_ms_per_step'], d['run_info']['isolated_stage_ms'])
"
