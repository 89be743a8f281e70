#!/bin/bash
TAG=${1:-r03k}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_posegraph.py tests/test_gpu_next.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "rc=$?" >> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value",round(d["value"]),"ms/step",round(d["ms_per_step"],2),"e2e",round(d["e2e"]["value"]),"dominant",d["roofline"]["kernel"],"frac",d["roofline"]["frac"],"traffic/alg",d["roofline"]["traffic_over_alg"])
print({k:round(v["ms_per_step"],2) for k,v in d["roofline"]["all_groups"].items()}, d["roofline"]["all_groups"]["local_ba_prv_windows"].get("summed_window_ms_per_step"))
PY
