#!/bin/bash
# round 2, step 8: the two copies of an x stage issued by different producer warps (nprod = 3)
set -u
O=gpurun_out
L=$O/r2_step8.log
: > $L
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3 >> $L
SH="256,512;256,256;128,256;64,128"
timeout 600 python tests/prof_sweep.py "$SH" "A3GC_TC_NPROD=1|A3GC_TC_NPROD=2|A3GC_TC_NPROD=3|A3GC_TC_NPROD=3 A3GC_TC_TRACE=1" 1024 40 fp32 A3GC >> $L 2>&1
timeout 600 python tests/prof_sweep.py "256,512" "A3GC_TC_NPROD=1|A3GC_TC_NPROD=3" 1024 40 fp32 AAGC >> $L 2>&1
timeout 600 python tests/prof_sweep.py "256,512" "A3GC_TC_NPROD=1|A3GC_TC_NPROD=3" 1024 40 bf16 A3GC >> $L 2>&1
tail -3 $L
