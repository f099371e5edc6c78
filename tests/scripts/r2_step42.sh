#!/bin/bash
# round 2, step 42: G-GRU kernel with the transposition buffers aliased into the state image (a 4th ring slot at H = 256)
set -u
O=gpurun_out
L=$O/r2_step42.log
: > $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $L
timeout 600 python tests/prof_sweep.py "256,512;256,256;128,256;64,128" "A3GC_TC_STAGES=3|A3GC_TC_OPT=0" 1024 40 fp32 GGRU >> $L 2>&1
timeout 600 python bench.py --variant GGRU --seq-len 600 --no-secondary --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400 >> $L
tail -3 $L
