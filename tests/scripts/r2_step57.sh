#!/bin/bash
# round 2, step 57: state exchange through L2 (bulk store + one multicast bulk load into the peers) instead of DSMEM copies
set -u
O=gpurun_out
L=$O/r2_step57.log
: > $L
K="A3GC_TC_OPT=0|A3GC_TC_XCHG=1 A3GC_TC_EARLYPUB=0|A3GC_TC_XCHG=1|A3GC_TC_EARLYPUB=0|A3GC_TC_XCHG=1 A3GC_TC_EARLYPUB=0 A3GC_TC_TRACE=1"
timeout 300 python tests/prof_sweep.py "256,512;256,256;128,256;128,128" "$K" 1024 40 fp32 A3GC >> $L 2>&1
echo "rc=$?" >> $L
timeout 200 python tests/prof_sweep.py "256,512" "A3GC_TC_OPT=0|A3GC_TC_XCHG=1 A3GC_TC_EARLYPUB=0" 1024 40 fp32 AAGC >> $L 2>&1
echo "rc=$?" >> $L
tail -3 $L | cut -c1-200
