#!/bin/bash
# round 2, step 21: backward chain -- L2 prefetch of the next iteration's tape rows (A3GC_BWD_PREFETCH)
set -u
O=gpurun_out
L=$O/r2_step21.log
: > $L
for pfv in 0 1; do
  echo "== A3GC_BWD_PREFETCH=$pfv" >> $L
  for shp in "256 12 3" "128 24 18" "64 15 9"; do
    A3GC_BWD_PREFETCH=$pfv A3GC_BWD_TRACE=1 timeout 600 python tests/prof_train.py $shp 256 200 2 2>&1 | grep -E "bwd trace|iter 2|bwd_blk" | tail -3 | sed 's/  *0.00%  *0.000us  *0.00%  *0.000us  *0.000us//' | cut -c1-230 >> $L
  done
done
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q 2>&1 | grep -E "rel_l2=|passed|failed" | head -8 >> $L
tail -5 $L
