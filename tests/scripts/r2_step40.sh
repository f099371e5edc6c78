#!/bin/bash
# round 2, step 40: two rings (weight slots + x-image slots), transposition buffers aliased into the state image
set -u
O=gpurun_out
L=$O/r2_step40.log
: > $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_round2.py tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -4 >> $L
SH="256,512;256,256;128,256;128,128;64,128;64,64"
timeout 900 python tests/prof_sweep.py "$SH" "A3GC_TC_STAGES=3|A3GC_TC_OPT=0|A3GC_TC_WSTAGES=3 A3GC_TC_XSTAGES=5|A3GC_TC_OPT=0 A3GC_TC_TRACE=1" 1024 40 fp32 A3GC >> $L 2>&1
timeout 600 python tests/prof_sweep.py "256,512;64,128" "A3GC_TC_OPT=0" 1024 40 fp32 AAGC >> $L 2>&1
timeout 600 python tests/prof_sweep.py "256,512;64,128" "A3GC_TC_OPT=0" 1024 40 bf16 A3GC >> $L 2>&1
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -3 >> $L
timeout 900 python bench.py --no-secondary --no-cpu-baseline 2>&1 | tail -1 >> $L
tail -3 $L
