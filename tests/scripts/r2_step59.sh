#!/bin/bash
# round 2, step 59: L2 state exchange in the training-mode forward as well (two all-gathers per step there)?
set -u
O=gpurun_out
L=$O/r2_step59.log
: > $L
A3GC_TC_XCHG=1 timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -2 >> $L
for x in 0 1; do
  echo "A3GC_TC_XCHG=$x" >> $L
  A3GC_TC_XCHG=$x A3GC_TC_TRACE=1 timeout 600 python tests/prof_train.py 256 12 3 256 200 2 2>&1 | grep -E "iter 2|step 4|steps 2" >> $L
done
tail -4 $L | cut -c1-250
