#!/bin/bash
# Round evidence in two gpurun calls (each bounded; gpurun_out/ must stay under 64 MiB to travel back):
#   bash tools/gpu_round.sh <tag> run      GPU parity tests, smoke, bench lines (ours / no-LBA / reference arm)
#   bash tools/gpu_round.sh <tag> profile  ncu launch list of the bench command + ncu --set full of the top kernels
# outputs under gpurun_out/<tag>_*
TAG=${1:-r}
WHAT=${2:-run}
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" >> gpurun_out/${TAG}_${WHAT}_steps.log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_nvsmi.txt
if [ "$WHAT" = run ]; then
  timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
  stamp pytest; tail -3 gpurun_out/${TAG}_pytest_gpu.log
  timeout 120 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log
  stamp smoke; tail -2 gpurun_out/${TAG}_smoke.log
  timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
  stamp bench; cat gpurun_out/${TAG}_bench.json
  timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
  stamp bench_ref
  timeout 200 python bench.py --steps 10 --warmup 3 --lba 0 --cpu-frames 8 > gpurun_out/${TAG}_bench_nolba.json 2> gpurun_out/${TAG}_bench_nolba.err
  stamp bench_nolba
else
  timeout 200 python -m pytest tests/test_gpu_next.py tests/test_gpu_orb.py -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu2.log
  stamp pytest2; tail -3 gpurun_out/${TAG}_pytest_gpu2.log
  timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench2.json 2> gpurun_out/${TAG}_bench2.err; echo "bench rc=$?" >> gpurun_out/${TAG}_bench2.err
  stamp bench2; cat gpurun_out/${TAG}_bench2.json
  timeout 330 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --cpu-frames 8 > gpurun_out/${TAG}_ncu_bench.log 2>&1
  stamp launches
  # ncu --set full in three short captures (one report each, summarised on the box; a report travels only if it is small):
  # the front-end kernels of one step with source, the tracking kernels, a slice of one LocalBA window's chain
  full() {  # name, kernel regex, skip, count, extra ncu flags, extra bench flags
    timeout 200 ncu --set full --clock-control none $5 -k regex:"$2" -s $3 -c $4 -o gpurun_out/${TAG}_prof_$1 \
      python bench.py --steps 1 --warmup 3 --cpu-frames 8 $6 > gpurun_out/${TAG}_ncu_full_$1.log 2>&1
    python tools/ncu_summary.py full gpurun_out/${TAG}_prof_$1.ncu-rep gpurun_out/${TAG}_ncu_full_$1.csv > /dev/null 2>&1
    SZ=$(stat -c %s gpurun_out/${TAG}_prof_$1.ncu-rep 2>/dev/null || echo 0)
    if [ "$SZ" -gt 20000000 ]; then rm -f gpurun_out/${TAG}_prof_$1.ncu-rep; echo "report $1 ($SZ bytes) summarised on the box and dropped" >> gpurun_out/${TAG}_${WHAT}_steps.log; fi
    stamp ncu_full_$1
  }
  full frontend 'k_fast_cells|k_pyramid|k_quadtree|k_orient_desc|k_stereo_match|k_stereo_filter' 6 6 "--import-source on" "--lba 0"
  full tracking 'k_sbp|k_frustum|k_pose_opt|k_imu_preint' 6 6 "" "--lba 0"
  full lba 'k_ba_' 120 24 "" "--lba-workers 1"
  python tools/ncu_summary.py launches gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_launches.md > /dev/null 2>&1
  stamp summaries; head -30 gpurun_out/${TAG}_launches.md
fi
cat gpurun_out/${TAG}_${WHAT}_steps.log
du -sh gpurun_out
