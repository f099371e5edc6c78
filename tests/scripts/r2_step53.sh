#!/bin/bash
# round 2, step 53: producer warps of the LSTM kernel on the whole warp as well (elected lane issues the bulk copies)
set -u
O=gpurun_out
L=$O/r2_step53.log
: > $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> $L
SH="256,512;256,256;128,256;128,128;64,128;64,64"
timeout 900 python tests/prof_sweep.py "$SH" "A3GC_TC_OPT=0|A3GC_TC_NPROD=2|A3GC_TC_NPROD=1" 1024 40 fp32 A3GC >> $L 2>&1
timeout 900 python tests/prof_sweep.py "256,512;64,128" "A3GC_TC_OPT=0" 1024 40 bf16 AAGC >> $L 2>&1
timeout 900 python bench.py --no-secondary --no-cpu-baseline 2>&1 | tail -1 >> $L
tail -3 $L | cut -c1-300
