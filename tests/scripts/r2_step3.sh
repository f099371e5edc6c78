#!/bin/bash
# round 2, step 3: multi-warp producers
set -u
O=gpurun_out
L=$O/r2_step3.log
: > $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $L
SH="256,512;256,256;128,256;128,128;64,128;64,64"
timeout 600 python tests/prof_sweep.py "$SH" "A3GC_TC_NPROD=1|A3GC_TC_NPROD=2|A3GC_TC_NPROD=3|A3GC_TC_NPROD=3 A3GC_TC_TRACE=1" >> $L 2>&1
timeout 300 python tests/prof_sweep.py "256,512;128,256;64,128" "A3GC_TC_NPROD=1|A3GC_TC_NPROD=3" 1024 40 bf16 >> $L 2>&1
timeout 300 python tests/prof_sweep.py "256,512;128,256" "A3GC_TC_NPROD=1|A3GC_TC_NPROD=3" 1024 40 fp32 AAGC >> $L 2>&1
timeout 300 python bench.py --no-cpu-baseline >> $L 2>&1
tail -3 $L
