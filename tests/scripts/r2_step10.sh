#!/bin/bash
# round 2, step 10: G-GRU kernel with the multi-warp producers
set -u
O=gpurun_out
L=$O/r2_step10.log
: > $L
timeout 900 python -m pytest tests -m gpu -x -q -k "GGRU or ggru or gru" 2>&1 | tail -3 >> $L
timeout 600 python tests/prof_sweep.py "256,512;256,256;128,256;64,128" "A3GC_TC_NPROD=1|A3GC_TC_NPROD=2|A3GC_TC_NPROD=3" 1024 120 fp32 GGRU >> $L 2>&1
tail -3 $L
