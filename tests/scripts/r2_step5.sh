#!/bin/bash
# round 2, step 5: training-step breakdown (torch profiler, stage 1 shape of cfg 5)
set -u
O=gpurun_out
L=$O/r2_step5.log
: > $L
timeout 600 python tests/prof_train.py 256 12 3 256 200 40 >> $L 2>&1
timeout 600 python tests/prof_train.py 128 15 9 256 200 25 >> $L 2>&1
tail -3 $L
