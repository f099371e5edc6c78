#!/bin/bash
# round 2, step 47: training-mode forward with per-step tape bases (immediate-offset stores in the gate phase)
set -u
O=gpurun_out
L=$O/r2_step47.log
: > $L
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -2 >> $L
for hh in "256 12 3" "64 15 3" "128 15 9"; do
  A3GC_TC_TRACE=1 timeout 600 python tests/prof_train.py $hh 256 200 2 2>&1 | grep -E "iter 2|step 4|steps 2" >> $L
done
timeout 900 python bench.py --workload train --no-cpu-baseline 2>&1 | tail -1 | cut -c1-900 >> $L
tail -3 $L
