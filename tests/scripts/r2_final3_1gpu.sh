#!/bin/bash
# round 2, final tree (two rings / aliased transposition buffers / training forward with per-step tape bases): GPU tests, smoke, the bench
# lines, side lines, phase traces, training-step breakdown, ncu launch list and the full capture of the graph-GRU kernel
set -u
O=gpurun_out
L=$O/r2_final3.log
: > $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1
timeout 900 python bench.py > $O/r02_bench.json 2> $O/r02_bench.err
timeout 600 python bench.py --impl reference > $O/r02_bench_reference.json 2>> $O/r02_bench.err
: > $O/r02_sidelines.jsonl
for vp in "A3GC bf16" "AAGC fp32" "AGC fp32" "GGRU fp32"; do
  set -- $vp
  timeout 600 python bench.py --variant $1 --precision $2 --no-cpu-baseline --no-secondary >> $O/r02_sidelines.jsonl 2>> $O/r02_bench.err
done
SH="256,512;256,256;128,256;128,128;64,128;64,64"
timeout 600 python tests/prof_sweep.py "$SH" "A3GC_TC_TRACE=1" > $O/r02_phase_traces.txt 2>&1
: > $O/r02_train_step_breakdown.txt
for shp in "256 12 3" "128 24 18" "64 15 9"; do
  A3GC_BWD_TRACE=1 timeout 600 python tests/prof_train.py $shp 256 200 30 2>&1 | grep -v -E "Warn|_warn_once|^$" | awk '!/bwd trace/ || !seen[$0]++' >> $O/r02_train_step_breakdown.txt
done
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --streams 1"
$CMD > $O/r02_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_ncu_launches.csv $CMD > $O/r02_ncu_launches.log 2>&1
CMDG="python bench.py --variant GGRU --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --streams 1"
$CMDG > $O/r02_ncu_plain_ggru.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gru_layer_kernel -s 19 -c 1 -o $O/r02b_ncu_ggru_f512_h256 -f $CMDG > $O/r02_ncu_full_ggru.log 2>&1
ncu -i $O/r02b_ncu_ggru_f512_h256.ncu-rep --page raw --csv > $O/r02b_ncu_raw_ggru_f512_h256.csv 2>> $L
ncu -i $O/r02b_ncu_ggru_f512_h256.ncu-rep --page details --csv > $O/r02b_ncu_details_ggru_f512_h256.csv 2>> $L
tail -3 $L
ls -la $O | tail -12 >> $L
