#!/bin/bash
# round 2, step 58: L2 state exchange on by default for 4-CTA clusters (inference): full GPU tests, sweep, bench
set -u
O=gpurun_out
L=$O/r2_step58.log
: > $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> $L
timeout 600 python tests/prof_sweep.py "256,512;256,256" "A3GC_TC_XCHG=0|A3GC_TC_OPT=0" 1024 40 fp32 AAGC >> $L 2>&1
timeout 600 python tests/prof_sweep.py "256,512;256,256" "A3GC_TC_XCHG=0|A3GC_TC_OPT=0" 1024 40 bf16 AAGC >> $L 2>&1
timeout 600 python tests/prof_sweep.py "256,512;256,256" "A3GC_TC_XCHG=0|A3GC_TC_OPT=0" 1024 40 bf16 A3GC >> $L 2>&1
timeout 600 python tests/prof_sweep.py "256,512;256,256" "A3GC_TC_XCHG=0|A3GC_TC_OPT=0" 1024 40 fp32 AGC >> $L 2>&1
timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -1 >> $L
tail -3 $L | cut -c1-300
