#!/bin/bash
# round 2, step 51: knob sweep on the kernel with the lean MMA issue loop (the old optima were found while the issue thread was the bottleneck)
set -u
O=gpurun_out
L=$O/r2_step51.log
: > $L
K="A3GC_TC_OPT=0|A3GC_TC_SPLIT=45,35|A3GC_TC_SPLIT=30,30|A3GC_TC_SPLIT=20,40|A3GC_TC_SPLIT=75,15|A3GC_TC_SPLIT=10,30|A3GC_TC_XPREFETCH=1|A3GC_TC_WSTAGES=3 A3GC_TC_XSTAGES=5|A3GC_TC_NPROD=2|A3GC_TC_NPROD=1|A3GC_TC_XSPLIT=1|A3GC_TC_XDEFER=1|A3GC_TC_EARLYPUB=0|A3GC_TC_ACOLL=0"
timeout 900 python tests/prof_sweep.py "256,512;256,256" "$K" 1024 40 fp32 A3GC >> $L 2>&1
K2="A3GC_TC_OPT=0|A3GC_TC_SPLIT=45,35|A3GC_TC_SPLIT=20,40|A3GC_TC_SPLIT=60,20|A3GC_TC_SPLIT=10,30|A3GC_TC_XPREFETCH=1|A3GC_TC_NPROD=1|A3GC_TC_XSPLIT=1"
timeout 900 python tests/prof_sweep.py "128,256;128,128;64,128;64,64" "$K2" 1024 40 fp32 A3GC >> $L 2>&1
tail -3 $L | cut -c1-200
