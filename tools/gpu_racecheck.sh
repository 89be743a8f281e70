#!/bin/bash
# compute-sanitizer racecheck / synccheck over the shared-memory pipelines written this round
TAG=${1:-r03j}
mkdir -p gpurun_out
run() { timeout $1 compute-sanitizer --tool $5 --print-limit 8 python -m pytest tests/$2 -m gpu -x -q -k "$3" > gpurun_out/${TAG}_$5_$4.log 2>&1; echo "$5 $4 rc=$?"; grep "SUMMARY\|passed\|failed\|hazard" gpurun_out/${TAG}_$5_$4.log | sort | uniq -c | sort -rn | head -6; }
run 200 test_gpu_ba.py "gba_single_step" gba racecheck
run 200 test_gpu_posegraph.py "single_damped_step" posegraph racecheck
run 120 test_gpu_ba.py "gba_single_step" gba synccheck
