#!/bin/bash
# round 2, step 11: byte-FIFO stream ring (variable-size stages)
set -u
O=gpurun_out
L=$O/r2_step11.log
: > $L
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $L
SH="256,512;256,256;128,256;128,128;64,128;64,64"
timeout 600 python tests/prof_sweep.py "$SH" "A3GC_TC_TRACE=1" 1024 40 fp32 A3GC >> $L 2>&1
timeout 600 python tests/prof_sweep.py "256,512;128,256" "A3GC_TC_OPT=0" 1024 40 fp32 AAGC >> $L 2>&1
timeout 600 python tests/prof_sweep.py "256,512;128,256" "A3GC_TC_OPT=0" 1024 40 bf16 A3GC >> $L 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-secondary >> $L 2>&1
tail -3 $L
