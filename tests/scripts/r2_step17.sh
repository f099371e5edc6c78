#!/bin/bash
# round 2, step 17: ncu --set full of the backward chain kernel (H=256, 3xTF32 contractions)
set -u
O=gpurun_out
L=$O/r2_step17.log
: > $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_train_bwd_blk -s 2 -c 1 -o $O/r02_bwd_blk_h256 -f \
  python tests/prof_train.py 256 12 3 256 200 2 > $O/r2_step17_ncu.log 2>&1
echo "ncu rc=$?" >> $L
ls -la $O/r02_bwd_blk_h256.ncu-rep >> $L 2>&1
tail -3 $L
