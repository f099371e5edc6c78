#!/bin/bash
# round 2, step 45: G-GRU kernel, sensitivity to the number of bulk-copy producer threads
set -u
O=gpurun_out
L=$O/r2_step45.log
: > $L
timeout 600 python tests/prof_sweep.py "256,512;256,256;128,256" "A3GC_TC_NPROD=1|A3GC_TC_NPROD=2|A3GC_TC_NPROD=3" 1024 40 fp32 GGRU >> $L 2>&1
tail -3 $L
