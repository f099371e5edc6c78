#!/bin/bash
# round 2, step 27: step 26 + 256-bit global accesses for the 64-byte tape rows
set -u
O=gpurun_out
L=$O/r2_step27.log
: > $L
for m in 4; do
  echo "== A3GC_BWD_MMA=$m" >> $L
  for shp in "256 12 3" "128 24 18" "64 15 9"; do
    A3GC_BWD_MMA=$m A3GC_BWD_TRACE=1 timeout 600 python tests/prof_train.py $shp 256 200 2 2>&1 | grep -E "bwd trace|iter 2|bwd_blk" | tail -3 | sed 's/  *0.00%  *0.000us  *0.00%  *0.000us  *0.000us//;s/\[a3gc bwd trace\] //' | cut -c1-230 >> $L
  done
  A3GC_BWD_MMA=$m timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q 2>&1 | grep -E "rel_l2=|passed|failed" | head -8 >> $L
done
timeout 600 python tests/diag_train_parity.py 2>&1 | grep -E "MMA=[24]" >> $L
tail -5 $L
timeout 900 python bench.py --workload train --no-cpu-baseline 2>&1 | tail -1 >> $L
