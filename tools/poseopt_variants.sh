#!/bin/bash
# k_pose_opt variants (tools/build_po_variant.sh -> build/po/libvieo_<v>.so): isolated time for 256 problems and parity
# (tests/test_gpu_ba.py pose tests) for the product library and each variant named on the command line; variants whose
# name starts with "prof" only print the in-kernel cycle profile.
mkdir -p gpurun_out
OUT=gpurun_out/${TAG:-po}_variants.txt
echo "== product" >> $OUT
timeout 120 python tools/poseopt_profile.py 128 >> $OUT 2>&1
timeout 200 python -m pytest tests/test_gpu_ba.py -m gpu -q -x -k "pose" 2>&1 | tail -3 >> $OUT
for v in "$@"; do
  echo "== $v" >> $OUT
  VIEO_B200_LIB=$PWD/build/po/libvieo_$v.so timeout 120 python tools/poseopt_profile.py 128 >> $OUT 2>&1
  case $v in prof*) ;; *) VIEO_B200_LIB=$PWD/build/po/libvieo_$v.so timeout 200 python -m pytest tests/test_gpu_ba.py -m gpu -q -x -k "pose" 2>&1 | tail -1 >> $OUT;; esac
done
cat $OUT
