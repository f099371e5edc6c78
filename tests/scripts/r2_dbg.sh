#!/bin/bash
set -u
O=gpurun_out
L=$O/r2_dbg.log
: > $L
for m in 0 1; do
  A3GC_BWD_TRACE=1 A3GC_BWD_MMA=$m timeout 600 python tests/prof_train.py 256 12 3 256 200 2 2>&1 | grep -E "bwd trace|iter 2" >> $L
  A3GC_BWD_TRACE=1 A3GC_BWD_MMA=$m timeout 600 python tests/prof_train.py 64 15 9 256 200 2 2>&1 | grep -E "bwd trace|iter 2" >> $L
done
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q 2>&1 | tail -2 >> $L
tail -5 $L
