#!/bin/bash
# round 2, step 44: ncu --set full of (a) the training-mode forward kernel (H=256, B=256 x T=200), (b) the inference kernel with the
# two rings (stage-1 rnn2 of the bench step)
set -u
O=gpurun_out
L=$O/r2_step44.log
: > $L
timeout 600 python tests/prof_train.py 256 12 3 256 200 2 > $O/r2_step44_plain.log 2>&1
echo "plain rc=$?" >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_lstm_layer_kernel -s 3 -c 1 -o $O/r02_ncu_trainfwd_h256 -f \
  python tests/prof_train.py 256 12 3 256 200 2 > $O/r2_step44_ncu.log 2>&1
echo "ncu train rc=$?" >> $L
ncu -i $O/r02_ncu_trainfwd_h256.ncu-rep --page raw --csv > $O/r02_ncu_raw_trainfwd_h256.csv 2>> $L
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --streams 1"
$CMD > $O/r2_step44_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_lstm_layer_kernel -s 19 -c 1 -o $O/r03_ncu_lstm_f512_h256 -f $CMD > $O/r2_step44_ncu2.log 2>&1
echo "ncu infer rc=$?" >> $L
ncu -i $O/r03_ncu_lstm_f512_h256.ncu-rep --page raw --csv > $O/r03_ncu_raw_lstm_f512_h256.csv 2>> $L
ncu -i $O/r03_ncu_lstm_f512_h256.ncu-rep --page details --csv > $O/r03_ncu_details_lstm_f512_h256.csv 2>> $L
ls -la $O | tail -12 >> $L
tail -3 $L
