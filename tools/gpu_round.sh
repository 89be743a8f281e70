#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench lines, ncu launch list, ncu --set full of the top kernels.
# usage: bash tools/gpu_round.sh [tag]   (outputs under gpurun_out/<tag>_*)
TAG=${1:-r}
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_nvsmi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
timeout 900 python bench.py --steps 20 --warmup 3 --lba 0 > gpurun_out/${TAG}_bench_nolba.json 2> gpurun_out/${TAG}_bench_nolba.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --cpu-frames 8 > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fast_cells|k_orient_desc|k_quadtree|k_resize|k_stereo_match|k_sbp|k_pose_opt|k_imu_preint|k_ba_linearize|k_ba_schur|k_ba_chol' -s 60 -c 36 -o gpurun_out/${TAG}_prof python bench.py --steps 1 --warmup 3 --cpu-frames 8 --lba-workers 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out
