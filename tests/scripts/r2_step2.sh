#!/bin/bash
# round 2, step 2: decoupled q chain (vector image) parity + phase traces
set -u
O=gpurun_out
L=$O/r2_step2.log
: > $L
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -s 2>&1 | tail -8 >> $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 >> $L
SH="256,512;256,256;128,256;128,128;64,128;64,64"
timeout 600 python tests/prof_sweep.py "$SH" "A3GC_TC_TRACE=1" >> $L 2>&1
timeout 300 python tests/prof_sweep.py "256,512;128,256;64,128" "A3GC_TC_TRACE=0" 1024 40 bf16 >> $L 2>&1
tail -5 $L
