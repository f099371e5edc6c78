#!/bin/bash
# round 2, step 20: phase trace + ncu --set full of the backward chain kernel (H=256, mode 2)
set -u
O=gpurun_out
L=$O/r2_step20.log
: > $L
A3GC_BWD_TRACE=1 timeout 600 python tests/prof_train.py 256 12 3 256 200 2 2>&1 | grep -E "bwd trace" | tail -2 >> $L
A3GC_BWD_TRACE=1 timeout 600 python tests/prof_train.py 64 15 9 256 200 2 2>&1 | grep -E "bwd trace" | tail -2 >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_train_bwd_blk -s 2 -c 1 -o $O/r02_bwd_blk_h256_v2 -f \
  python tests/prof_train.py 256 12 3 256 200 2 > $O/r2_step20_ncu.log 2>&1
echo "ncu rc=$?" >> $L
tail -5 $L
