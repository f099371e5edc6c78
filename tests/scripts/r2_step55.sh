#!/bin/bash
# round 2, step 55: ncu --set full of the dominant launch on the FINAL tree (lean issue loop, x-image prefetch)
set -u
O=gpurun_out
L=$O/r2_step55.log
: > $L
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --streams 1"
$CMD > $O/r2_step55_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_lstm_layer_kernel -s 19 -c 1 -o $O/r02c_ncu_lstm_f512_h256 -f $CMD > $O/r2_step55_ncu.log 2>&1
echo "ncu rc=$?" >> $L
ncu -i $O/r02c_ncu_lstm_f512_h256.ncu-rep --page raw --csv > $O/r02c_ncu_raw_lstm_f512_h256.csv 2>> $L
ncu -i $O/r02c_ncu_lstm_f512_h256.ncu-rep --page details --csv > $O/r02c_ncu_details_lstm_f512_h256.csv 2>> $L
tail -2 $L
