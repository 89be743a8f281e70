#!/bin/bash
# guarded GBA iteration: one small parity test under a short timeout first (a pipeline deadlock must not burn the budget)
TAG=${1:-r03d}
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_ba.py -m gpu -x -q -k "gba_single_step" > gpurun_out/${TAG}_pytest_quick.log 2>&1; RC=$?
echo "quick rc=$RC" >> gpurun_out/${TAG}_pytest_quick.log; tail -3 gpurun_out/${TAG}_pytest_quick.log
if [ $RC -ne 0 ]; then exit 0; fi
timeout 240 python -m pytest tests/test_gpu_ba.py -m gpu -x -q -k "gba or global" > gpurun_out/${TAG}_pytest_gba.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gba.log
tail -3 gpurun_out/${TAG}_pytest_gba.log
timeout 120 python bench.py --config 4 --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_c4_n1.json 2> gpurun_out/${TAG}_bench_c4.err; echo "rc=$?" >> gpurun_out/${TAG}_bench_c4.err
cut -c1-260 gpurun_out/${TAG}_bench_c4_n1.json
GBA_ORACLE=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/${TAG}_launches_gba.csv \
  python tools/gba_profile.py 400 25000 20 > gpurun_out/${TAG}_ncu_gba.log 2>&1
python tools/ncu_summary.py launches gpurun_out/${TAG}_launches_gba.csv gpurun_out/${TAG}_launches_gba.md > /dev/null 2>&1
rm -f gpurun_out/${TAG}_launches_gba.csv
head -14 gpurun_out/${TAG}_launches_gba.md
