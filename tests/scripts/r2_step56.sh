#!/bin/bash
# round 2, step 56: last knob sweep on the final tree (x-image prefetch on): ring split and x-part segments at H = 256
set -u
O=gpurun_out
L=$O/r2_step56.log
: > $L
K="A3GC_TC_OPT=0|A3GC_TC_WSTAGES=3 A3GC_TC_XSTAGES=5|A3GC_TC_SPLIT=60,20|A3GC_TC_SPLIT=50,30|A3GC_TC_SPLIT=40,40|A3GC_TC_SPLIT=35,45|A3GC_TC_SPLIT=25,50|A3GC_TC_SPLIT=55,35|A3GC_TC_NPROD=2|A3GC_TC_OPT=0"
timeout 900 python tests/prof_sweep.py "256,512;256,256" "$K" 1024 40 fp32 A3GC >> $L 2>&1
tail -3 $L | cut -c1-200
