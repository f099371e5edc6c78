#!/bin/bash
# round 2, step 41: how much the G-GRU kernel depends on the ring depth (is the H = 256 kernel with 3 slots ring-bound?)
set -u
O=gpurun_out
L=$O/r2_step41.log
: > $L
timeout 600 python tests/prof_sweep.py "256,512;256,256" "A3GC_TC_STAGES=2|A3GC_TC_STAGES=3" 1024 40 fp32 GGRU >> $L 2>&1
timeout 600 python tests/prof_sweep.py "128,256;128,128" "A3GC_TC_STAGES=2|A3GC_TC_STAGES=3|A3GC_TC_STAGES=4|A3GC_TC_STAGES=6" 1024 40 fp32 GGRU >> $L 2>&1
timeout 600 python tests/prof_sweep.py "128,256" "A3GC_TC_STAGES=2|A3GC_TC_STAGES=3|A3GC_TC_STAGES=4|A3GC_TC_STAGES=6" 1024 40 fp32 A3GC >> $L 2>&1
tail -3 $L
