#!/bin/bash
# round 2, step 18: backward chain after the two-instruction TF32 split (modes 1 = 3xTF32, 2 = TF32 head + bf16 corrections)
set -u
O=gpurun_out
L=$O/r2_step18.log
: > $L
for m in 1 2; do
  echo "== A3GC_BWD_MMA=$m" >> $L
  A3GC_BWD_TRACE=1 A3GC_BWD_MMA=$m timeout 600 python tests/prof_train.py 256 12 3 256 200 2 2>&1 | grep -E "bwd trace|iter 2|bwd_blk" | tail -3 | cut -c1-200 >> $L
  A3GC_BWD_TRACE=1 A3GC_BWD_MMA=$m timeout 600 python tests/prof_train.py 128 24 18 256 200 2 2>&1 | grep -E "bwd trace|iter 2|bwd_blk" | tail -3 | cut -c1-200 >> $L
  A3GC_BWD_TRACE=1 A3GC_BWD_MMA=$m timeout 600 python tests/prof_train.py 64 15 9 256 200 2 2>&1 | grep -E "bwd trace|iter 2|bwd_blk" | tail -3 | cut -c1-200 >> $L
  A3GC_BWD_MMA=$m timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q 2>&1 | grep -E "rel_l2=|passed|failed" | head -8 >> $L
done
timeout 600 python tests/diag_train_parity.py >> $L 2>&1
tail -5 $L
