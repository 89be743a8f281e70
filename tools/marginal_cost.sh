#!/bin/bash
# Marginal cost of each kernel group of the bench step (diagnostic): the step with kernel groups left out
# (VIEO_BENCH_SKIP), optionally with a variant library.   usage: TAG=.. LBA=0 bash tools/marginal_cost.sh "<skip list>" ...
# an argument "lib=<name>" switches the following runs to build/po/libvieo_<name>.so ("lib=" back to the product)
mkdir -p gpurun_out
OUT=gpurun_out/${TAG:-mc}_marginal.txt
LIBV=""
for sk in "$@"; do
  case $sk in lib=*) LIBV=${sk#lib=}; continue;; esac
  [ "$sk" = none ] && sk=""
  if [ -n "$LIBV" ]; then export VIEO_B200_LIB=$PWD/build/po/libvieo_$LIBV.so; else unset VIEO_B200_LIB; fi
  VIEO_BENCH_SKIP=$sk timeout 120 python bench.py --steps 10 --warmup 3 --lba ${LBA:-0} --cpu-frames 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('lib=%-8s skip=%-14s ms/step %.2f  e2e %.0f  groups %s' % ('$LIBV', '$sk', d['ms_per_step'], d['e2e']['value'], {k[:12]: round(v['ms_per_step'],2) for k,v in d['roofline']['all_groups'].items()}))" >> $OUT 2>&1
done
cat $OUT
