#!/bin/bash
# round 2, step 7: A-operand collector reuse (UTCHMMA .A_KEEP / .A_REUSE) on / off
set -u
O=gpurun_out
L=$O/r2_step7.log
: > $L
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $L
SH="256,512;256,256;128,256;64,128"
timeout 600 python tests/prof_sweep.py "$SH" "A3GC_TC_ACOLL=0|A3GC_TC_ACOLL=1" 1024 40 fp32 A3GC >> $L 2>&1
timeout 600 python tests/prof_sweep.py "$SH" "A3GC_TC_ACOLL=0|A3GC_TC_ACOLL=1" 1024 120 fp32 GGRU >> $L 2>&1
timeout 600 python tests/prof_sweep.py "256,512" "A3GC_TC_ACOLL=0|A3GC_TC_ACOLL=1" 1024 40 fp32 AAGC >> $L 2>&1
tail -3 $L
