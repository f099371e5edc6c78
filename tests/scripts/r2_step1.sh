#!/bin/bash
# round 2, step 1: HSEP (separate recurrent accumulator) parity + phase traces + x-split sweep
set -u
O=gpurun_out
L=$O/r2_step1.log
: > $L
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 >> $L
SH="256,512;256,256;128,256;128,128;64,128;64,64"
timeout 600 python tests/prof_sweep.py "$SH" "A3GC_TC_HSEP=0 A3GC_TC_TRACE=1|A3GC_TC_HSEP=1 A3GC_TC_TRACE=1" >> $L 2>&1
SP=""
for sp in "0,0" "0,50" "0,100" "25,25" "25,50" "40,30" "50,25" "50,50" "60,40" "75,25" "100,0" "30,70"; do SP="$SP|A3GC_TC_SPLIT=$sp"; done
timeout 900 python tests/prof_sweep.py "$SH" "A3GC_TC_HSEP=0$SP" >> $L 2>&1
timeout 300 python tests/prof_sweep.py "256,512;128,256;64,128" "A3GC_TC_HSEP=0|A3GC_TC_HSEP=1" 1024 40 bf16 >> $L 2>&1
tail -5 $L
