#!/bin/bash
# k_pose_opt register / CTA-size variants (built into build/po/ by hand): isolated time for 256 problems and
# parity (tests/test_gpu_ba.py pose tests) for each.
mkdir -p gpurun_out
for v in base lb2 t128 t128lb3 t128lb4; do
  echo "== $v" >> gpurun_out/po_variants.txt
  VIEO_B200_LIB=$PWD/build/po/libvieo_$v.so timeout 120 python tools/poseopt_profile.py 128 >> gpurun_out/po_variants.txt 2>&1
  VIEO_B200_LIB=$PWD/build/po/libvieo_$v.so timeout 200 python -m pytest tests/test_gpu_ba.py -m gpu -q -x -k "pose" 2>&1 | tail -1 >> gpurun_out/po_variants.txt
done
echo "== prof" >> gpurun_out/po_variants.txt
VIEO_B200_LIB=$PWD/build/po/libvieo_prof.so timeout 120 python tools/poseopt_profile.py 128 >> gpurun_out/po_variants.txt 2>&1
cat gpurun_out/po_variants.txt
