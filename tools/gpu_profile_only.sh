#!/bin/bash
# ncu evidence of the bench command: launch list (gpu__time_duration.sum) + --set full summaries of the front-end / tracking kernels
TAG=${1:-r03g}
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" >> gpurun_out/${TAG}_profile_steps.log; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 3 --cpu-frames 8 > gpurun_out/${TAG}_ncu_bench.log 2>&1
stamp launches
python tools/ncu_summary.py launches gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_launches.md > /dev/null 2>&1
rm -f gpurun_out/${TAG}_launches.csv
full() {
  timeout 200 ncu --set full --clock-control none $5 -k regex:"$2" -s $3 -c $4 -o gpurun_out/${TAG}_prof_$1 \
    python bench.py --steps 1 --warmup 3 --cpu-frames 8 $6 > gpurun_out/${TAG}_ncu_full_$1.log 2>&1
  python tools/ncu_summary.py full gpurun_out/${TAG}_prof_$1.ncu-rep gpurun_out/${TAG}_ncu_full_$1.csv > /dev/null 2>&1
  rm -f gpurun_out/${TAG}_prof_$1.ncu-rep
  stamp ncu_full_$1
}
full frontend 'k_fast_cells|k_pyramid|k_quadtree|k_orient_desc|k_stereo_match|k_stereo_filter' 6 6 "" "--lba 0"
full tracking 'k_sbp|k_frustum|k_pose_opt|k_imu_preint' 6 8 "" "--lba 0"
head -28 gpurun_out/${TAG}_launches.md
cat gpurun_out/${TAG}_profile_steps.log
