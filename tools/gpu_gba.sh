#!/bin/bash
# global-BA iteration on the GPU: parity tests of the BA file, configs[4] bench line, ncu launch list of one solve
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ba.py -m gpu -x -q -k "gba or global" > gpurun_out/${TAG}_pytest_gba.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gba.log
tail -5 gpurun_out/${TAG}_pytest_gba.log
timeout 300 python bench.py --config 4 --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_c4_n1.json 2> gpurun_out/${TAG}_bench_c4.err; echo "rc=$?" >> gpurun_out/${TAG}_bench_c4.err
cat gpurun_out/${TAG}_bench_c4_n1.json | cut -c1-700
if [ "$2" = prof ]; then
  GBA_ORACLE=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/${TAG}_launches_gba.csv \
    python tools/gba_profile.py 400 25000 20 > gpurun_out/${TAG}_ncu_gba.log 2>&1
  python tools/ncu_summary.py launches gpurun_out/${TAG}_launches_gba.csv gpurun_out/${TAG}_launches_gba.md > /dev/null 2>&1
  rm -f gpurun_out/${TAG}_launches_gba.csv
  head -32 gpurun_out/${TAG}_launches_gba.md
fi
