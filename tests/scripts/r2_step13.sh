#!/bin/bash
# round 2, step 13: separate ring for the attention weights (A3GC_TC_TSLOTS)
set -u
O=gpurun_out
L=$O/r2_step13.log
: > $L
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $L
SET="A3GC_TC_TSLOTS=0|A3GC_TC_TSLOTS=2|A3GC_TC_TSLOTS=3|A3GC_TC_TSLOTS=4"
timeout 600 python tests/prof_sweep.py "128,256;128,128;64,128;64,64" "$SET" 1024 40 fp32 A3GC >> $L 2>&1
timeout 600 python tests/prof_sweep.py "256,512;256,256;128,256;64,128" "$SET" 1024 40 bf16 A3GC >> $L 2>&1
timeout 600 python tests/prof_sweep.py "256,512;128,256" "$SET" 1024 40 bf16 AGC >> $L 2>&1
timeout 300 python tests/prof_sweep.py "256,512;128,256" "A3GC_TC_TSLOTS=4 A3GC_TC_TRACE=1" 1024 40 bf16 A3GC >> $L 2>&1
tail -3 $L
