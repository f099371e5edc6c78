#!/bin/bash
# round 2, step 16: backward chain -- TF32 head + bf16 corrections (A3GC_BWD_MMA=2), 16 loads in flight in the GEMVs
set -u
O=gpurun_out
L=$O/r2_step16.log
: > $L
for m in "1 0" "1 1" "2 1"; do
  set -- $m
  echo "== A3GC_BWD_MMA=$1 A3GC_BWD_GEMV16=$2" >> $L
  A3GC_BWD_TRACE=1 A3GC_BWD_MMA=$1 A3GC_BWD_GEMV16=$2 timeout 600 python tests/prof_train.py 256 12 3 256 200 2 2>&1 | grep -E "bwd trace|iter 2|bwd_blk" | tail -4 | cut -c1-60,150-250 >> $L
  A3GC_BWD_TRACE=1 A3GC_BWD_MMA=$1 A3GC_BWD_GEMV16=$2 timeout 600 python tests/prof_train.py 128 24 18 256 200 2 2>&1 | grep -E "bwd trace|iter 2|bwd_blk" | tail -4 | cut -c1-60,150-250 >> $L
  A3GC_BWD_MMA=$1 A3GC_BWD_GEMV16=$2 timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q 2>&1 | grep -E "rel_l2=|passed|failed" | head -8 >> $L
done
tail -5 $L
