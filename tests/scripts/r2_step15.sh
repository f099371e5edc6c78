#!/bin/bash
# round 2, step 15: backward chain with the two weight contractions on the tensor cores (3xTF32 mma.sync), A3GC_BWD_MMA
set -u
O=gpurun_out
L=$O/r2_step15.log
: > $L
timeout 1200 python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -6 >> $L
for m in 0 1; do
  echo "== A3GC_BWD_MMA=$m" >> $L
  A3GC_BWD_MMA=$m timeout 600 python tests/prof_train.py 256 12 3 256 200 6 2>&1 | grep -v Warn >> $L
  A3GC_BWD_MMA=$m timeout 600 python tests/prof_train.py 64 15 9 256 200 4 2>&1 | grep -v Warn >> $L
  A3GC_BWD_MMA=$m timeout 600 python tests/prof_train.py 128 24 18 256 200 4 2>&1 | grep -v Warn >> $L
done
timeout 900 python bench.py --workload train --no-cpu-baseline 2>&1 | tail -1 >> $L
tail -3 $L
