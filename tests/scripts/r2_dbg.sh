#!/bin/bash
set -u
O=gpurun_out
L=$O/r2_dbg.log
: > $L
for shp in "256 12 3" "64 15 9"; do
  A3GC_TC_TRACE=1 timeout 600 python tests/prof_train.py $shp 256 200 2 2>&1 | grep -E "step|iter 2" >> $L
done
timeout 300 python tests/prof_sweep.py "256,512;64,128" "A3GC_TC_TRACE=1" 256 200 fp32 A3GC >> $L 2>&1
tail -5 $L
