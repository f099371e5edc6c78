#!/bin/bash
# round 2, step 43: phase trace of the G-GRU kernel
set -u
O=gpurun_out
L=$O/r2_step43.log
: > $L
for s in "256 512" "256 256" "128 256" "64 128"; do timeout 300 python tests/prof_gru_trace.py $s 2>&1 | grep -v Warning >> $L; done
tail -3 $L
