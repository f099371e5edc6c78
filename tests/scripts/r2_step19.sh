#!/bin/bash
# round 2, step 19: backward chain -- batched loads in phase A, mask prefetch in G; default mode 2
set -u
O=gpurun_out
L=$O/r2_step19.log
: > $L
A3GC_BWD_TRACE=1 timeout 600 python tests/prof_train.py 256 12 3 256 200 8 2>&1 | grep -E "bwd trace|iter 2|bwd_blk|gemm|sgemm|tc_lstm" | tail -9 | cut -c1-60,150-250 >> $L
A3GC_BWD_TRACE=1 timeout 600 python tests/prof_train.py 128 24 18 256 200 2 2>&1 | grep -E "bwd trace|iter 2|bwd_blk" | tail -3 | cut -c1-60,150-250 >> $L
A3GC_BWD_TRACE=1 timeout 600 python tests/prof_train.py 64 15 9 256 200 2 2>&1 | grep -E "bwd trace|iter 2|bwd_blk" | tail -3 | cut -c1-60,150-250 >> $L
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q 2>&1 | grep -E "rel_l2=|passed|failed" | head -8 >> $L
timeout 600 python tests/diag_train_parity.py 2>&1 | grep "MMA=2" >> $L
timeout 900 python bench.py --workload train --no-cpu-baseline 2>&1 | tail -1 >> $L
tail -5 $L
