#!/bin/bash
# round 2, after the training-path work: GPU tests, smoke, the bench line, the backward kernel's phase trace and ncu capture
set -u
O=gpurun_out
L=$O/r2_final2.log
: > $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1
timeout 900 python bench.py > $O/r02_bench.json 2> $O/r02_bench.err
for shp in "256 12 3" "128 24 18" "64 15 9"; do
  A3GC_BWD_TRACE=1 timeout 600 python tests/prof_train.py $shp 256 200 30 2>&1 | grep -v -E "Warn|_warn_once|^$" | awk '!/bwd trace/ || !seen[$0]++' >> $O/r02_train_step_breakdown.txt
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_train_bwd_blk -s 2 -c 1 -o $O/r02_bwd_blk_h256_final -f \
  python tests/prof_train.py 256 12 3 256 200 2 > $O/r2_final2_ncu.log 2>&1
echo "ncu rc=$?" >> $L
tail -3 $L
