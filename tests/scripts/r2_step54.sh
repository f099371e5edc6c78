#!/bin/bash
# round 2, step 54: G-GRU kernel with the L2 prefetch of the next step's x image
set -u
O=gpurun_out
L=$O/r2_step54.log
: > $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_fullsize.py -m gpu -x -q -k "GGRU or ggru or gru or smoke or chain" 2>&1 | tail -2 >> $L
timeout 600 python tests/prof_sweep.py "256,512;256,256;128,256;128,128;64,128;64,64" "A3GC_TC_XPREFETCH=0|A3GC_TC_XPREFETCH=1" 1024 40 fp32 GGRU >> $L 2>&1
timeout 600 python bench.py --variant GGRU --seq-len 600 --no-secondary --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200 >> $L
tail -3 $L
