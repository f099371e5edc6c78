#!/bin/bash
# round 2, step 14: bf16 local rescale of the peers' blocks instead of the second all-gather (A3GC_TC_RESCALE_BF16)
set -u
O=gpurun_out
L=$O/r2_step14.log
: > $L
SET="A3GC_TC_RESCALE_BF16=0|A3GC_TC_RESCALE_BF16=1"
timeout 600 python tests/diag_bf16_error.py "$SET" >> $L 2>&1
timeout 600 python tests/prof_sweep.py "256,512;256,256;128,256;128,128" "$SET" 1024 40 bf16 A3GC >> $L 2>&1
timeout 600 python tests/prof_sweep.py "256,512;128,256" "$SET" 1024 40 bf16 AGC >> $L 2>&1
A3GC_TC_RESCALE_BF16=1 timeout 600 python -m pytest tests -m gpu -x -q -k "bf16" 2>&1 | tail -4 >> $L
tail -3 $L
