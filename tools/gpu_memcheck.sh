#!/bin/bash
# compute-sanitizer memcheck over the kernels written this round (small parity cases)
TAG=${1:-r03i}
mkdir -p gpurun_out
S="compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5"
run() { timeout $1 $S python -m pytest tests/$2 -m gpu -x -q -k "$3" > gpurun_out/${TAG}_memcheck_$4.log 2>&1; echo "$4 rc=$?"; grep -c "Invalid\|ERROR SUMMARY" gpurun_out/${TAG}_memcheck_$4.log; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/${TAG}_memcheck_$4.log | tail -3; }
run 150 test_gpu_posegraph.py "single_damped_step or isolated or correct_points" posegraph
run 150 test_gpu_ba.py "gba_single_step or gba_scale_single_step" gba
run 100 test_gpu_next.py "frustum_rig" rig
