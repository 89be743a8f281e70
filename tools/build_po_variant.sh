#!/bin/bash
# build/po/libvieo_<name>.so = the product library with poseopt.cu recompiled with extra -D flags:
#   bash tools/build_po_variant.sh <name> "<flags>"      (used with VIEO_B200_LIB=... by tools/poseopt_variants.sh)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/po
N=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC $* -c vieo_slam_b200/csrc/poseopt.cu -o build/po/poseopt_$N.o
OBJS=$(ls vieo_slam_b200/csrc/*.o | grep -v poseopt.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/po/libvieo_$N.so $OBJS build/po/poseopt_$N.o
echo built build/po/libvieo_$N.so
