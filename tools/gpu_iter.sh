#!/bin/bash
# one development iteration on the GPU: BA parity tests, the default bench line, the configs[4] line
TAG=${1:-r02y}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ba.py tests/test_sharding.py -m gpu -x -q > gpurun_out/${TAG}_pytest_ba.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_ba.log
tail -4 gpurun_out/${TAG}_pytest_ba.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "rc=$?" >> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("value",round(d["value"]),"ms/step",round(d["ms_per_step"],2),"e2e",round(d["e2e"]["value"]),"lat",d.get("single_frame_latency_ms",{}).get("median"),"lba_window_ms",d["run_info"]["isolated_stage_ms"])
    print({k:round(v["ms_per_step"],2) for k,v in d["roofline"]["all_groups"].items()})
except Exception as e: print("bench parse failed",e); print(open("gpurun_out/${TAG}_bench.err").read()[-1500:])
PY
timeout 300 python bench.py --config 4 --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_c4_n1.json 2> gpurun_out/${TAG}_bench_c4.err; echo "rc=$?" >> gpurun_out/${TAG}_bench_c4.err
cut -c1-260 gpurun_out/${TAG}_bench_c4_n1.json
