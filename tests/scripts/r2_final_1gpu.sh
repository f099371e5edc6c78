#!/bin/bash
# round 2: final single-GPU artefacts (tests, smoke, bench lines, ncu launch list + full capture of the dominant launch)
set -u
O=gpurun_out
L=$O/r2_final.log
: > $L
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1
timeout 900 python bench.py > $O/r02_bench.json 2> $O/r02_bench.err
timeout 600 python bench.py --impl reference > $O/r02_bench_reference.json 2>> $O/r02_bench.err
for vp in "A3GC bf16" "AAGC fp32" "AGC fp32" "GGRU fp32"; do
  set -- $vp
  timeout 600 python bench.py --variant $1 --precision $2 --no-cpu-baseline --no-secondary >> $O/r02_sidelines.jsonl 2>> $O/r02_bench.err
done
SH="256,512;256,256;128,256;128,128;64,128;64,64"
timeout 600 python tests/prof_sweep.py "$SH" "A3GC_TC_TRACE=1" > $O/r02_phase_traces.txt 2>&1
# ncu: launch list of one bench step (warm-up launches skipped), then the full set on the dominant launch (stage-1 rnn2)
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --streams 1"
$CMD > $O/r02_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_ncu_launches.csv $CMD > $O/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_lstm_layer_kernel -s 19 -c 1 -o $O/r02_ncu_lstm_f512_h256 $CMD > $O/r02_ncu_full.log 2>&1
# the restructured graph-GRU kernel: full capture of its dominant launch (stage-1 rnn2 of G-GRU-TP)
CMDG="python bench.py --variant GGRU --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --streams 1"
$CMDG > $O/r02_ncu_plain_ggru.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_gru_layer_kernel -s 19 -c 1 -o $O/r02_ncu_ggru_f512_h256 $CMDG > $O/r02_ncu_full_ggru.log 2>&1
tail -3 $L
ls -la $O | tail -12 >> $L
